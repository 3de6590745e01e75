"""mannequin/_utils.py equivalents: endswith, discounted (device scan), bar."""
import numpy as np
import torch

from . import _device as D


def endswith(a, b):
    """True when sequence `a` ends with `b` (mannequin/_utils.py:3-7)."""
    if len(a) < len(b):
        return False
    return a[len(a) - len(b):] == b


def discounted(values, *, horizon):
    """Normalised reverse EMA of rewards (mannequin/_utils.py:9-20), fp64 like the
    reference, computed by the reverse-scan kernel (csrc/mq_scan.cu)."""
    step = 1.0 / float(horizon)
    assert (step > 1e-6) and (step < 1.0)
    values = np.array(values, dtype=np.float64)
    n = len(values)
    if n == 0:
        return values
    x = D.from_host(values.reshape(-1), np.float64)
    out = torch.empty_like(x)
    wsb = D._lib.lib().mq_scan_ws_bytes(n)
    ws = D.workspace("scan", wsb)
    D.call("mq_discounted", D.ptr(x), n, float(horizon), D.ptr(out), D.ptr(ws), wsb, D.stream())
    return D.to_host(out).reshape(values.shape)


def bar(value, max_value=100.0, *, length=50):
    """Text progress bar (mannequin/_utils.py:22-46); host-only, non-numeric."""
    value = float(value)
    max_value = abs(float(max_value))
    length = max(1, int(length))
    frac_total = max(0.0, min(1.0, abs(value) / max_value))
    if value >= 0.0:
        full, part = divmod(int(round(frac_total * length * 8)), 8)
        body = chr(0x2588) * full + (chr(0x2590 - part) if part >= 1 else "")
        body = (body + " " * length)[:length]
    else:
        full, part = divmod(int(round(frac_total * length * 2)), 2)
        body = (chr(0x257A) if part >= 1 else "") + chr(0x2501) * full
        body = (" " * length + body)[-length:]
    width = len("-%.2f" % max_value)
    text = ("%%%d.2f " % width) % value
    return text[0:width + 1] + "[" + body + "]"
