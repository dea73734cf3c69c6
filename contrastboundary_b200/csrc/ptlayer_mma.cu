// ptlayer_mma.cu — the dense contraction of the PointTransformer local aggregation on the tensor cores.
// Reference: linear_w[2] of PointTransformerLayer (pytorch/model/blocks.py:24-28,40): for every (point, neighbour)
// row, w2 = W3 relu(bn2(w0)) + b3 — the only genuine (n*k) x c x c/8 dense contraction of the layer.
//
// One warp owns a tile of 16 consecutive rows of the (n*k) row space (one neighbourhood for k = 16) = one m16 MMA
// tile.  The A operand u = relu(bn2(x_k[idx] - x_q + pr)) never exists in memory: every lane gathers and computes
// exactly the elements of its A fragment.  The contraction index (channel) is PERMUTED inside each group of 8
// channels — virtual column t <-> channel 2t, t+4 <-> channel 2t+1 — so that a lane's fragment elements are two
// adjacent channels of rows g and g+8; two k-steps are processed together, so a lane gathers 16 bytes per row and
// the four lanes of a quad read 64 contiguous bytes of a feature row.  B = W3 is staged once per CTA in shared memory (split into TF32 hi / lo)
// with the same permutation folded into its indexing.  3xTF32 (tf32.cuh) keeps FP32-level accuracy.
#include "ptlayer.cuh"
#include "tf32.cuh"

#define PM_THREADS 256
#define PM_WARPS (PM_THREADS / 32)

static int g_pt_mma = 1;
extern "C" int cb_pt_set_tensor_cores(int on) { g_pt_mma = on ? 1 : 0; return g_pt_mma; }
int cb_pt_mma_enabled() { return g_pt_mma; }

struct PmSmall { float w1[9], b1[3], sc1[3], sh1[3]; };
__device__ __forceinline__ PmSmall pm_small_load(const float *__restrict__ d)
{
    PmSmall s;
#pragma unroll
    for (int i = 0; i < 9; i++) s.w1[i] = __ldg(d + i);
#pragma unroll
    for (int i = 0; i < 3; i++) { s.b1[i] = __ldg(d + 9 + i); s.sc1[i] = __ldg(d + 12 + i); s.sh1[i] = __ldg(d + 15 + i); }
    return s;
}
__device__ __forceinline__ void pm_g1(const PmSmall &sp, float rx, float ry, float rz, float (&g)[3])
{
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float h = sp.w1[3 * a] * rx + sp.w1[3 * a + 1] * ry + sp.w1[3 * a + 2] * rz + sp.b1[a];
        g[a] = fmaxf(h * sp.sc1[a] + sp.sh1[a], 0.f);
    }
}

// shared-memory size of k_pt_w2_mma<C>
template <int C> struct PmW2 {
    static constexpr int CS = C / 8;
    static constexpr int NT = (CS + 7) / 8;          // n-tiles of 8 output columns
    static constexpr int CSP = NT * 8;               // padded output columns (rows of W3 beyond CS are zero)
    static constexpr int LDW = C + 16;               // row stride of the staged W3: (LDW / 4) mod 8 == 4 -> conflict-free LDS.128
    static constexpr bool PRESPLIT = C <= 256;       // hi and lo copies fit in shared memory
    static constexpr size_t smem = (size_t)(PRESPLIT ? 2 : 1) * CSP * LDW * 4 + (size_t)C * 8 * 4 + 2 * CS * 4;
};

template <int C>
__global__ void __launch_bounds__(PM_THREADS) k_pt_w2_mma(int n, int k, int ld, const float *__restrict__ rel,
                                                          const int *__restrict__ idx, const float *__restrict__ xq,
                                                          const float *__restrict__ xk, const float *__restrict__ w2p,
                                                          const float *__restrict__ b2p, const float *__restrict__ smalld,
                                                          const float *__restrict__ bn2 /* [4][C] */,
                                                          const float *__restrict__ w3 /* [CS][C] */,
                                                          const float *__restrict__ b3, float *__restrict__ w2out,
                                                          double *__restrict__ stats /* [2][CS] */,
                                                          const float *__restrict__ w0 /* (n,k,C) or NULL */)
{
    using T = PmW2<C>;
    constexpr int CS = T::CS, NT = T::NT, CSP = T::CSP, LDW = T::LDW;
    extern __shared__ __align__(16) unsigned char pm_sm[];
    unsigned *Wh = reinterpret_cast<unsigned *>(pm_sm);                       // [CSP][LDW]  tf32 hi (or raw fp32 when !PRESPLIT)
    unsigned *Wl = Wh + (T::PRESPLIT ? CSP * LDW : 0);                        // [CSP][LDW]  tf32 lo
    float *P = reinterpret_cast<float *>(Wh + (T::PRESPLIT ? 2 : 1) * CSP * LDW);   // [C][8]: w2a w2b w2c b2 sc2 sh2 - -
    float *comb = P + C * 8;                                                  // [2][CS]
    const PmSmall sp = pm_small_load(smalld);
    for (int e = threadIdx.x; e < CSP * C; e += PM_THREADS) {
        const int i = e / C, c = e % C;
        const float v = i < CS ? __ldg(w3 + (size_t)i * C + c) : 0.f;
        if (T::PRESPLIT) {
            unsigned hi, lo;
            tg_split(v, hi, lo);
            Wh[i * LDW + c] = hi; Wl[i * LDW + c] = lo;
        } else {
            Wh[i * LDW + c] = __float_as_uint(v);
        }
    }
    for (int c = threadIdx.x; c < C; c += PM_THREADS) {
        float *p = P + c * 8;
        p[0] = __ldg(w2p + 3 * c); p[1] = __ldg(w2p + 3 * c + 1); p[2] = __ldg(w2p + 3 * c + 2); p[3] = __ldg(b2p + c);
        p[4] = __ldg(bn2 + c); p[5] = __ldg(bn2 + C + c); p[6] = 0.f; p[7] = 0.f;
    }
    for (int i = threadIdx.x; i < 2 * CS; i += PM_THREADS) comb[i] = 0.f;
    __syncthreads();

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const long long rows = (long long)n * k;
    const long long tiles = (rows + 15) >> 4;
    float t1[NT][2], t2[NT][2], bias3[NT][2];
#pragma unroll
    for (int jn = 0; jn < NT; jn++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            t1[jn][e] = 0.f; t2[jn][e] = 0.f;
            const int col = jn * 8 + 2 * t + e;
            bias3[jn][e] = col < CS ? __ldg(b3 + col) : 0.f;
        }
    for (long long tile = (long long)blockIdx.x * PM_WARPS + wib; tile < tiles; tile += (long long)gridDim.x * PM_WARPS) {
        const long long rA = tile * 16 + g, rB = rA + 8;
        const bool vA = rA < rows, vB = rB < rows;
        const long long ra = vA ? rA : 0, rb = vB ? rB : 0;
        const float *xkA = xk + (size_t)__ldg(idx + ra) * ld, *xkB = xk + (size_t)__ldg(idx + rb) * ld;
        const float *xqA = xq + (size_t)(ra / k) * ld, *xqB = xq + (size_t)(rb / k) * ld;
        float gA[3], gB[3];
        pm_g1(sp, __ldg(rel + 3 * ra), __ldg(rel + 3 * ra + 1), __ldg(rel + 3 * ra + 2), gA);
        pm_g1(sp, __ldg(rel + 3 * rb), __ldg(rel + 3 * rb + 1), __ldg(rel + 3 * rb + 2), gB);
        float acc[NT][4];
#pragma unroll
        for (int jn = 0; jn < NT; jn++) acc[jn][0] = acc[jn][1] = acc[jn][2] = acc[jn][3] = 0.f;
#pragma unroll 2
        for (int j = 0; j < C / 16; j++) {
            // 16 channels = two k-steps; this lane owns channels ch .. ch+3 of rows g and g+8
            const int ch = 16 * j + 4 * t;
            float uA[4], uB[4];
            if (w0) {
                // training: w0 was materialised by the statistics pass (it is kept for the backward) — stream it
                const float4 wa4 = __ldg(reinterpret_cast<const float4 *>(w0 + ra * C + ch)), wb4 = __ldg(reinterpret_cast<const float4 *>(w0 + rb * C + ch));
                const float wav[4] = {wa4.x, wa4.y, wa4.z, wa4.w}, wbv[4] = {wb4.x, wb4.y, wb4.z, wb4.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float2 pb = *reinterpret_cast<const float2 *>(P + (ch + e) * 8 + 4);
                    uA[e] = vA ? fmaxf(wav[e] * pb.x + pb.y, 0.f) : 0.f;
                    uB[e] = vB ? fmaxf(wbv[e] * pb.x + pb.y, 0.f) : 0.f;
                }
            } else {
                const float4 xa = __ldg(reinterpret_cast<const float4 *>(xkA + ch)), xb = __ldg(reinterpret_cast<const float4 *>(xkB + ch));
                const float4 qa = __ldg(reinterpret_cast<const float4 *>(xqA + ch)), qb = __ldg(reinterpret_cast<const float4 *>(xqB + ch));
                const float xav[4] = {xa.x, xa.y, xa.z, xa.w}, xbv[4] = {xb.x, xb.y, xb.z, xb.w};
                const float qav[4] = {qa.x, qa.y, qa.z, qa.w}, qbv[4] = {qb.x, qb.y, qb.z, qb.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float4 p = *reinterpret_cast<const float4 *>(P + (ch + e) * 8), pb = *reinterpret_cast<const float4 *>(P + (ch + e) * 8 + 4);
                    // u = relu(bn2(x_k - x_q + W2 g1 + b2)), same operation order as the SIMT kernels
                    const float prA = p.x * gA[0] + p.y * gA[1] + p.z * gA[2] + p.w;
                    const float prB = p.x * gB[0] + p.y * gB[1] + p.z * gB[2] + p.w;
                    uA[e] = vA ? fmaxf((xav[e] - qav[e] + prA) * pb.x + pb.y, 0.f) : 0.f;
                    uB[e] = vB ? fmaxf((xbv[e] - qbv[e] + prB) * pb.x + pb.y, 0.f) : 0.f;
                }
            }
            unsigned ah[2][4], al[2][4];
#pragma unroll
            for (int s2 = 0; s2 < 2; s2++) {
                tg_split(uA[2 * s2], ah[s2][0], al[s2][0]);          // (row g,   virtual k t)   = channel ch + 2 s2
                tg_split(uB[2 * s2], ah[s2][1], al[s2][1]);          // (row g+8, virtual k t)
                tg_split(uA[2 * s2 + 1], ah[s2][2], al[s2][2]);      // (row g,   virtual k t+4) = channel ch + 2 s2 + 1
                tg_split(uB[2 * s2 + 1], ah[s2][3], al[s2][3]);      // (row g+8, virtual k t+4)
            }
#pragma unroll
            for (int jn = 0; jn < NT; jn++) {
                const int off = (jn * 8 + g) * LDW + ch;      // B(k, n = g): W3[8 jn + g][channel]
                unsigned bh[4], bl[4];
                if (T::PRESPLIT) {
                    const uint4 h = *reinterpret_cast<const uint4 *>(Wh + off), l = *reinterpret_cast<const uint4 *>(Wl + off);
                    bh[0] = h.x; bh[1] = h.y; bh[2] = h.z; bh[3] = h.w; bl[0] = l.x; bl[1] = l.y; bl[2] = l.z; bl[3] = l.w;
                } else {
                    const uint4 w = *reinterpret_cast<const uint4 *>(Wh + off);
                    tg_split(__uint_as_float(w.x), bh[0], bl[0]);
                    tg_split(__uint_as_float(w.y), bh[1], bl[1]);
                    tg_split(__uint_as_float(w.z), bh[2], bl[2]);
                    tg_split(__uint_as_float(w.w), bh[3], bl[3]);
                }
#pragma unroll
                for (int s2 = 0; s2 < 2; s2++) {
                    tg_mma(acc[jn], al[s2], bh[2 * s2], bh[2 * s2 + 1]);
                    tg_mma(acc[jn], ah[s2], bl[2 * s2], bl[2 * s2 + 1]);
                    tg_mma(acc[jn], ah[s2], bh[2 * s2], bh[2 * s2 + 1]);
                }
            }
        }
#pragma unroll
        for (int jn = 0; jn < NT; jn++) {
            const int col = jn * 8 + 2 * t;
            if (col < CS) {
                const float oA0 = acc[jn][0] + bias3[jn][0], oA1 = acc[jn][1] + bias3[jn][1];
                const float oB0 = acc[jn][2] + bias3[jn][0], oB1 = acc[jn][3] + bias3[jn][1];
                if (vA) {
                    *reinterpret_cast<float2 *>(w2out + rA * CS + col) = make_float2(oA0, oA1);
                    t1[jn][0] += oA0; t2[jn][0] += oA0 * oA0; t1[jn][1] += oA1; t2[jn][1] += oA1 * oA1;
                }
                if (vB) {
                    *reinterpret_cast<float2 *>(w2out + rB * CS + col) = make_float2(oB0, oB1);
                    t1[jn][0] += oB0; t2[jn][0] += oB0 * oB0; t1[jn][1] += oB1; t2[jn][1] += oB1 * oB1;
                }
            }
        }
    }
    // BatchNorm statistics of w2: reduce over the 8 row-lanes (g), then block, then global (double)
#pragma unroll
    for (int jn = 0; jn < NT; jn++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            float a = t1[jn][e], b = t2[jn][e];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                a += __shfl_xor_sync(CB_FULL_MASK, a, o);
                b += __shfl_xor_sync(CB_FULL_MASK, b, o);
            }
            const int col = jn * 8 + 2 * t + e;
            if (g == 0 && col < CS) { atomicAdd(&comb[col], a); atomicAdd(&comb[CS + col], b); }
        }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * CS; i += PM_THREADS) atomicAdd(stats + i, (double)comb[i]);
}

template <int C>
static void pm_w2_launch(int n, int k, int ld, const float *rel, const int *idx, const float *xq, const float *xk,
                         const float *w2p, const float *b2p, const float *smalld, const float *bn2, const float *w3,
                         const float *b3, float *w2out, double *stats, const float *w0, cudaStream_t st)
{
    const size_t smem = PmW2<C>::smem;
    cudaFuncSetAttribute(k_pt_w2_mma<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long tiles = ((long long)n * k + 15) / 16;
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    long long blocks = (tiles + PM_WARPS - 1) / PM_WARPS;
    if (blocks > 148LL * per_sm) blocks = 148LL * per_sm;
    if (blocks < 1) blocks = 1;
    k_pt_w2_mma<C><<<(int)blocks, PM_THREADS, smem, st>>>(n, k, ld, rel, idx, xq, xk, w2p, b2p, smalld, bn2, w3, b3, w2out, stats, w0);
}

// dispatch used by ptlayer_fwd.cu
void cb_pt_w2_mma(int c, int n, int k, int ld, const float *rel, const int *idx, const float *xq, const float *xk,
                  const float *w2p, const float *b2p, const float *smalld, const float *bn2, const float *w3, const float *b3,
                  float *w2out, double *stats, const float *w0, cudaStream_t st)
{
    switch (c) {
    case 32: pm_w2_launch<32>(n, k, ld, rel, idx, xq, xk, w2p, b2p, smalld, bn2, w3, b3, w2out, stats, w0, st); break;
    case 64: pm_w2_launch<64>(n, k, ld, rel, idx, xq, xk, w2p, b2p, smalld, bn2, w3, b3, w2out, stats, w0, st); break;
    case 128: pm_w2_launch<128>(n, k, ld, rel, idx, xq, xk, w2p, b2p, smalld, bn2, w3, b3, w2out, stats, w0, st); break;
    case 256: pm_w2_launch<256>(n, k, ld, rel, idx, xq, xk, w2p, b2p, smalld, bn2, w3, b3, w2out, stats, w0, st); break;
    case 512: pm_w2_launch<512>(n, k, ld, rel, idx, xq, xk, w2p, b2p, smalld, bn2, w3, b3, w2out, stats, w0, st); break;
    default: break;
    }
}
