// Developer aid (not part of the library): probes the tcgen05.mma kind::f16 (bf16) operand layouts the
// tensor-core MLP kernel relies on.  Operands are staged in shared memory as 128-byte core
// matrices (8 x 16 B, SWIZZLE_NONE); each case names a shared-memory image, the descriptor
// strides and the per-K-step start-address advance, and the accumulator tile is dumped
// lane by lane from TMEM and compared with a host product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tools/tc_probe.cu && ./tc_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

struct Case {
    uint32_t a_lbo, a_sbo, a_step;   // bytes
    uint32_t b_lbo, b_sbo, b_step;   // bytes
    uint32_t a_major, b_major;       // 0 = K-major, 1 = MN-major
    uint32_t m, n, ksteps;
    uint32_t a_bytes, b_bytes;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
    return d;                        // layout type 0 = SWIZZLE_NONE, base offset 0
}

__global__ void __launch_bounds__(128) probe_kernel(Case c, const uint16_t *a_img, const uint16_t *b_img, float *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t mbar;
    uint16_t *sa = (uint16_t *)smem;
    uint16_t *sb = (uint16_t *)(smem + c.a_bytes);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < c.a_bytes / 2; i += 128) sa[i] = a_img[i];
    for (uint32_t i = tid; i < c.b_bytes / 2; i += 128) sb[i] = b_img[i];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = tmem_base;
    // clear the accumulator region (lanes x 256 columns) so untouched lanes read as a marker
    {
        const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
        for (int col = 0; col < 256; col += 8) {
            uint32_t v = __float_as_uint(-777.0f);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr + col), "r"(v));
        }
        asm volatile("tcgen05.wait::st.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (tid == 0) {
        uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (c.a_major << 15) | (c.b_major << 16) |
                         ((c.n >> 3) << 17) | ((c.m >> 4) << 24);
        for (uint32_t k = 0; k < c.ksteps; ++k) {
            uint64_t da = make_desc(smem_u32(sa) + k * c.a_step, c.a_lbo, c.a_sbo);
            uint64_t db = make_desc(smem_u32(sb) + k * c.b_step, c.b_lbo, c.b_sbo);
            uint32_t acc = k > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    {
        const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
        for (int col = 0; col < 256; col += 8) {
            uint32_t v[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(taddr + col));
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            for (int j = 0; j < 8; ++j) out[tid * 256 + col + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u));
}

// host side ---------------------------------------------------------------------------------
static uint16_t bf16_bits(float x) { uint32_t u; memcpy(&u, &x, 4); return (uint16_t)(u >> 16); }
static float bf16_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFF0000u; memcpy(&x, &u, 4); return x; }

// image of a logical [rows][cols] matrix as 8 x (16 B = 8 bf16) core matrices: chunk (rg, fc) at rg*rg_stride + fc*fc_stride
static std::vector<uint16_t> image(const std::vector<float> &mat, int rows, int cols, int rg_stride, int fc_stride, int bytes) {
    std::vector<uint16_t> img(bytes / 2, 0);
    for (int r = 0; r < rows; ++r)
        for (int f = 0; f < cols; ++f) {
            int off = (r / 8) * rg_stride + (f / 8) * fc_stride + (r % 8) * 16 + (f % 8) * 2;
            img[off / 2] = bf16_bits(mat[(size_t)r * cols + f]);
        }
    return img;
}

static int run_case(const char *name, Case c, const std::vector<uint16_t> &a_img, const std::vector<uint16_t> &b_img,
                    const std::vector<float> &want /* m x n */) {
    uint16_t *da, *db; float *dout;
    cudaMalloc(&da, c.a_bytes); cudaMalloc(&db, c.b_bytes); cudaMalloc(&dout, 128 * 256 * 4);
    cudaMemcpy(da, a_img.data(), c.a_bytes, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b_img.data(), c.b_bytes, cudaMemcpyHostToDevice);
    size_t smem = c.a_bytes + c.b_bytes;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<<<1, 128, smem>>>(c, da, db, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return 1; }
    std::vector<float> out(128 * 256);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    const char *hyp[2] = {"lane=i", "lane=i%16+32*(i/16)"};
    for (int h = 0; h < 2; ++h) {
        double maxerr = 0, maxref = 0;
        for (uint32_t i = 0; i < c.m; ++i)
            for (uint32_t j = 0; j < c.n; ++j) {
                int lane = h == 0 ? i : (i % 16) + 32 * (i / 16);
                if (lane >= 128) { maxerr = 1e30; continue; }
                double d = fabs((double)out[lane * 256 + j] - want[i * c.n + j]);
                if (d > maxerr) maxerr = d;
                if (fabs(want[i * c.n + j]) > maxref) maxref = fabs(want[i * c.n + j]);
            }
        printf("%-34s hyp[%s]: max err %.3e (max |ref| %.3e)\n", name, hyp[h], maxerr, maxref);
    }
    int touched_lanes = 0, first = -1, last = -1, maxcol = -1;
    for (int l = 0; l < 128; ++l) {
        bool t = false;
        for (int j = 0; j < 256; ++j) if (out[l * 256 + j] != -777.0f) { t = true; if (j > maxcol) maxcol = j; }
        if (t) { touched_lanes++; if (first < 0) first = l; last = l; }
    }
    printf("    touched lanes %d (first %d last %d), max column %d; out[0][0..3] = %g %g %g %g want %g %g %g %g\n",
           touched_lanes, first, last, maxcol, out[0], out[1], out[2], out[3], want[0], want[1], want[2], want[3]);
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return 0;
}

int main() {
    srand(1);
    auto rnd = [](int n) { std::vector<float> v(n); for (auto &x : v) x = bf16_trunc((float)rand() / RAND_MAX * 2 - 1); return v; };
    const int R = 128, F = 64;
    std::vector<float> H = rnd(R * F), G = rnd(R * F), W = rnd(F * F);
    // activation image: fc_stride = R*16 (2048: all rows of one 8-feature chunk), rg_stride = 128
    const int FCS = R * 16, RGS = 128, ABYTES = R * F * 2;
    auto Himg = image(H, R, F, RGS, FCS, ABYTES), Gimg = image(G, R, F, RGS, FCS, ABYTES);
    // weight image: Wt[n=out][k=in] K-major: fc_stride = 64*16 = 1024, rg_stride 128
    std::vector<float> Wt(F * F);
    for (int i = 0; i < F; ++i) for (int o = 0; o < F; ++o) Wt[o * F + i] = W[i * F + o];
    const int WFCS = F * 16, WBYTES = F * F * 2;
    auto Wimg = image(Wt, F, F, RGS, WFCS, WBYTES);
    {   // Z = H W  (A K-major rows x in, B K-major out x in), K step = 16 features = 2 chunks
        std::vector<float> want(R * F, 0.f);
        for (int r = 0; r < R; ++r) for (int o = 0; o < F; ++o) { double s = 0; for (int i = 0; i < F; ++i) s += (double)H[r * F + i] * W[i * F + o]; want[r * F + o] = (float)s; }
        Case c = {(uint32_t)FCS, (uint32_t)RGS, (uint32_t)(2 * FCS), (uint32_t)WFCS, (uint32_t)RGS, (uint32_t)(2 * WFCS), 0, 0, 128, 64, 4, (uint32_t)ABYTES, (uint32_t)WBYTES};
        run_case("fwd  A:K  B:K  M128 N64", c, Himg, Wimg, want);
        std::vector<float> want16(R * 16);
        for (int r = 0; r < R; ++r) for (int o = 0; o < 16; ++o) want16[r * 16 + o] = want[r * F + o];
        c.n = 16;
        run_case("fwd  A:K  B:K  M128 N16", c, Himg, Wimg, want16);
    }
    {   // dX = G W^T: B'[n'=in][k'=out] = same Wimg viewed MN-major: MN chunks (in/8) stride WFCS, K groups (out/8) stride RGS;
        // K step = 16 out = two row groups: +2*RGS
        std::vector<float> want(R * F, 0.f);
        for (int r = 0; r < R; ++r) for (int i = 0; i < F; ++i) { double s = 0; for (int o = 0; o < F; ++o) s += (double)G[r * F + o] * W[i * F + o]; want[r * F + i] = (float)s; }
        Case c = {(uint32_t)FCS, (uint32_t)RGS, (uint32_t)(2 * FCS), (uint32_t)RGS, (uint32_t)WFCS, (uint32_t)(2 * RGS), 0, 1, 128, 64, 4, (uint32_t)ABYTES, (uint32_t)WBYTES};
        run_case("dX   A:K  B:MN lbo=K sbo=MN", c, Gimg, Wimg, want);
        Case c2 = c; c2.b_lbo = WFCS; c2.b_sbo = RGS;
        run_case("dX   A:K  B:MN lbo=MN sbo=K", c2, Gimg, Wimg, want);
    }
    {   // dW = H^T G: A = H viewed MN-major (M = features, K = rows), B = G viewed MN-major (N = features, K = rows)
        std::vector<float> want(F * F, 0.f);
        for (int i = 0; i < F; ++i) for (int o = 0; o < F; ++o) { double s = 0; for (int r = 0; r < R; ++r) s += (double)H[r * F + i] * G[r * F + o]; want[i * F + o] = (float)s; }
        Case c = {(uint32_t)RGS, (uint32_t)FCS, (uint32_t)(2 * RGS), (uint32_t)RGS, (uint32_t)FCS, (uint32_t)(2 * RGS), 1, 1, 64, 64, 8, (uint32_t)ABYTES, (uint32_t)ABYTES};
        run_case("dW   A:MN B:MN M64 lbo=K sbo=MN", c, Himg, Gimg, want);
        Case c2 = c; c2.a_lbo = FCS; c2.a_sbo = RGS; c2.b_lbo = FCS; c2.b_sbo = RGS;
        run_case("dW   A:MN B:MN M64 lbo=MN sbo=K", c2, Himg, Gimg, want);
        // N = 8 (bias trick) and N = 16
        std::vector<float> want8(F * 8), want16(F * 16);
        for (int i = 0; i < F; ++i) { for (int o = 0; o < 8; ++o) want8[i * 8 + o] = want[i * F + o]; for (int o = 0; o < 16; ++o) want16[i * 16 + o] = want[i * F + o]; }
        Case c4 = c; c4.n = 8;
        run_case("dW   A:MN B:MN M64 N8", c4, Himg, Gimg, want8);
        Case c5 = c; c5.n = 16;
        run_case("dW   A:MN B:MN M64 N16", c5, Himg, Gimg, want16);
        // stacked M = 128: A = [H ; H2] with H2's image right behind H's (feature chunks 8..15)
        std::vector<float> H2 = rnd(R * F);
        auto H2img = image(H2, R, F, RGS, FCS, ABYTES);
        std::vector<uint16_t> stack(Himg); stack.insert(stack.end(), H2img.begin(), H2img.end());
        std::vector<float> want2(2 * F * F, 0.f);
        for (int i = 0; i < F; ++i) for (int o = 0; o < F; ++o) { double s = 0; for (int r = 0; r < R; ++r) s += (double)H2[r * F + i] * G[r * F + o]; want2[(F + i) * F + o] = (float)s; want2[i * F + o] = want[i * F + o]; }
        Case c3 = c; c3.m = 128; c3.a_bytes = 2 * ABYTES;
        run_case("dW   A:MN B:MN M128 stacked", c3, stack, Gimg, want2);
    }
    return 0;
}
