// linear_ops.cu — tall-skinny FP32 linear layers for the per-point MLPs around the local
// aggregation (reference: the nn.Linear calls of pytorch/model/blocks.py:33,72,76,108,127-131 and
// heads.py MLPs).  Their shapes are n x {6..160} x {13..128} with n up to 655 360 rows: 0.1-1.5 GFLOP,
// i.e. HBM-bound streaming problems, for which the library SGEMM heuristics pick 64x64 SIMT tiles and
// a split-K "largek" kernel that run at a few % of bandwidth.  Exact FP32 FMA accumulation (the 1e-4
// parity budget rules out TF32).
//   forward : Y[n,co]  = X[n,ci] W[co,ci]^T + b
//   dgrad   : dX[n,ci] = G[n,co] W[co,ci]
//   wgrad   : dW[co,ci] = G^T X ,  db[co] = sum_rows G
#include "common.cuh"

#define LG_BM 128      // rows per block tile
#define LG_BK 32       // k chunk
#define LG_THREADS 256
#define LG_APAD 132    // padded row-tile stride of the transposed A chunk (keeps LDS.128 aligned, spreads banks)

// C[n x N] = A[n x K] * B[K x N] (+ bias[N]);  B[k][j] = transB ? W[j*ldw + k] : W[k*ldw + j].
// One block: LG_BM rows x NT columns (NT = 16 * CPT); blockIdx.y tiles N.
template <int CPT>
__global__ void __launch_bounds__(LG_THREADS) k_skinny_gemm(int n, int K, int N, const float *__restrict__ A,
                                                            const float *__restrict__ W, int ldw, int transB,
                                                            const float *__restrict__ bias, float *__restrict__ C)
{
    constexpr int NT = 16 * CPT;
    __shared__ __align__(16) float As[LG_BK][LG_APAD];
    __shared__ __align__(16) float Bs[LG_BK][NT];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long row0 = (long long)blockIdx.x * LG_BM;
    const int col0 = blockIdx.y * NT;
    float acc[8][CPT];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int j = 0; j < CPT; j++) acc[r][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += LG_BK) {
        // A chunk: LG_BM x LG_BK, stored transposed
        for (int e = tid; e < LG_BM * LG_BK; e += LG_THREADS) {
            const int r = e / LG_BK, kk = e % LG_BK;
            const long long row = row0 + r;
            float v = 0.f;
            if (row < n && k0 + kk < K) v = __ldg(A + row * K + k0 + kk);
            As[kk][r] = v;
        }
        for (int e = tid; e < LG_BK * NT; e += LG_THREADS) {
            const int kk = e / NT, j = e % NT;
            float v = 0.f;
            if (k0 + kk < K && col0 + j < N)
                v = transB ? __ldg(W + (size_t)(col0 + j) * ldw + k0 + kk) : __ldg(W + (size_t)(k0 + kk) * ldw + col0 + j);
            Bs[kk][j] = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < LG_BK; kk++) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8 + 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[CPT];
#pragma unroll
            for (int j = 0; j < CPT; j++) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int j = 0; j < CPT; j++) acc[r][j] = fmaf(a[r], b[j], acc[r][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const long long row = row0 + ty * 8 + r;
        if (row >= n) continue;
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            const int col = col0 + tx + 16 * j;
            if (col < N) C[row * N + col] = acc[r][j] + (bias ? __ldg(bias + col) : 0.f);
        }
    }
}

static int launch_skinny(int n, int K, int N, const float *A, const float *W, int ldw, int transB, const float *bias,
                         float *C, cudaStream_t st)
{
    const int gx = (n + LG_BM - 1) / LG_BM;
    // column tile: the width in {128, 64, 32, 16} with the fewest padded columns (ties -> wider)
    int best = 128, waste = (N + 127) / 128 * 128 - N;
    for (int nt = 64; nt >= 16; nt >>= 1) {
        const int w = (N + nt - 1) / nt * nt - N;
        if (w < waste) { waste = w; best = nt; }
    }
    const dim3 grid(gx, (N + best - 1) / best);
    if (best == 16) k_skinny_gemm<1><<<grid, LG_THREADS, 0, st>>>(n, K, N, A, W, ldw, transB, bias, C);
    else if (best == 32) k_skinny_gemm<2><<<grid, LG_THREADS, 0, st>>>(n, K, N, A, W, ldw, transB, bias, C);
    else if (best == 64) k_skinny_gemm<4><<<grid, LG_THREADS, 0, st>>>(n, K, N, A, W, ldw, transB, bias, C);
    else k_skinny_gemm<8><<<grid, LG_THREADS, 0, st>>>(n, K, N, A, W, ldw, transB, bias, C);
    return 0;
}

// wgrad: dW[co][ci] += sum_rows G[row][co] * X[row][ci]; db[co] += sum_rows G[row][co]
// Block = chunk of rows; thread owns outputs e = tid + 256*t (t < OPT), e -> (o = e / ci, i = e % ci).
#define WG_ROWS 64
template <int OPT>
__global__ void __launch_bounds__(LG_THREADS) k_skinny_wgrad(int n, int ci, int co, const float *__restrict__ X,
                                                             const float *__restrict__ G, float *__restrict__ dW,
                                                             float *__restrict__ db, int rows_per_block)
{
    extern __shared__ float wsm[];
    float *Xs = wsm;                       // [WG_ROWS][ci]
    float *Gs = wsm + WG_ROWS * ci;        // [WG_ROWS][co]
    const int tid = threadIdx.x;
    float acc[OPT];
    int oo[OPT], ii[OPT];
#pragma unroll
    for (int t = 0; t < OPT; t++) {
        acc[t] = 0.f;
        const int e = tid + LG_THREADS * t;
        oo[t] = e < ci * co ? e / ci : -1;
        ii[t] = e < ci * co ? e % ci : 0;
    }
    float accb = 0.f;
    const long long r_begin = (long long)blockIdx.x * rows_per_block;
    long long r_end = r_begin + rows_per_block;
    if (r_end > n) r_end = n;
    for (long long r0 = r_begin; r0 < r_end; r0 += WG_ROWS) {
        const int rows = (int)((r_end - r0) < WG_ROWS ? (r_end - r0) : WG_ROWS);
        for (int e = tid; e < rows * ci; e += LG_THREADS) Xs[e] = __ldg(X + r0 * ci + e);
        for (int e = tid; e < rows * co; e += LG_THREADS) Gs[e] = __ldg(G + r0 * co + e);
        __syncthreads();
#pragma unroll
        for (int t = 0; t < OPT; t++) {
            if (oo[t] >= 0) {
                float s = 0.f;
                for (int r = 0; r < rows; r++) s = fmaf(Gs[r * co + oo[t]], Xs[r * ci + ii[t]], s);
                acc[t] += s;
            }
        }
        if (db && tid < co) {
            float s = 0.f;
            for (int r = 0; r < rows; r++) s += Gs[r * co + tid];
            accb += s;
        }
        __syncthreads();
    }
#pragma unroll
    for (int t = 0; t < OPT; t++)
        if (oo[t] >= 0) atomicAdd(dW + tid + LG_THREADS * t, acc[t]);
    if (db && tid < co) atomicAdd(db + tid, accb);
}

// register-tiled wgrad for ci % 4 == 0 and co % 4 == 0: thread owns 4x4 patches of dW
// (patch p -> outputs o in [4*(p / (ci/4)), +4), inputs i in [4*(p % (ci/4)), +4)); per staged row two LDS.128 feed 16 FMAs.
template <int PPT>
__global__ void __launch_bounds__(LG_THREADS) k_skinny_wgrad4(int n, int ci, int co, const float *__restrict__ X,
                                                              const float *__restrict__ G, float *__restrict__ dW,
                                                              float *__restrict__ db, int rows_per_block)
{
    extern __shared__ __align__(16) float wsm[];
    float *Xs = wsm;                       // [WG_ROWS][ci]
    float *Gs = wsm + WG_ROWS * ci;        // [WG_ROWS][co]
    const int tid = threadIdx.x;
    const int pci = ci >> 2, npatch = (co >> 2) * pci;
    float acc[PPT][4][4];
    int po[PPT], pi[PPT];
#pragma unroll
    for (int t = 0; t < PPT; t++) {
        const int p = tid + LG_THREADS * t;
        po[t] = p < npatch ? (p / pci) * 4 : -1;
        pi[t] = p < npatch ? (p % pci) * 4 : 0;
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[t][a][b] = 0.f;
    }
    float accb = 0.f;
    const long long r_begin = (long long)blockIdx.x * rows_per_block;
    long long r_end = r_begin + rows_per_block;
    if (r_end > n) r_end = n;
    for (long long r0 = r_begin; r0 < r_end; r0 += WG_ROWS) {
        const int rows = (int)((r_end - r0) < WG_ROWS ? (r_end - r0) : WG_ROWS);
        {
            const float4 *xs = reinterpret_cast<const float4 *>(X + r0 * ci);
            const float4 *gs = reinterpret_cast<const float4 *>(G + r0 * co);
            for (int e = tid; e < rows * ci / 4; e += LG_THREADS) reinterpret_cast<float4 *>(Xs)[e] = __ldg(xs + e);
            for (int e = tid; e < rows * co / 4; e += LG_THREADS) reinterpret_cast<float4 *>(Gs)[e] = __ldg(gs + e);
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < PPT; t++) {
            if (po[t] >= 0) {
                for (int r = 0; r < rows; r++) {
                    const float4 g = *reinterpret_cast<const float4 *>(Gs + r * co + po[t]);
                    const float4 x = *reinterpret_cast<const float4 *>(Xs + r * ci + pi[t]);
                    const float gg[4] = {g.x, g.y, g.z, g.w}, xx[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int b = 0; b < 4; b++) acc[t][a][b] = fmaf(gg[a], xx[b], acc[t][a][b]);
                }
            }
        }
        if (db && tid < co) {
            float s = 0.f;
            for (int r = 0; r < rows; r++) s += Gs[r * co + tid];
            accb += s;
        }
        __syncthreads();
    }
#pragma unroll
    for (int t = 0; t < PPT; t++)
        if (po[t] >= 0) {
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) atomicAdd(dW + (size_t)(po[t] + a) * ci + pi[t] + b, acc[t][a][b]);
        }
    if (db && tid < co) atomicAdd(db + tid, accb);
}

// tensor-core (3xTF32) variants, tc_gemm.cu
int cb_tc_enabled();
void cb_tc_linear_forward(int n, int ci, int co, const float *X, const float *W, const float *b, float *Y, cudaStream_t st);
void cb_tc_linear_dgrad(int n, int ci, int co, const float *G, const float *W, float *dX, cudaStream_t st);
void cb_tc_linear_wgrad(int n, int ci, int co, const float *X, const float *G, float *dW, float *db, const float *xsc,
                        const float *xsh, cudaStream_t st);

bool cb_umma_shape_ok(int n, int K, int N, const float *A, const float *Y, int lda, int ldy);
int cb_umma_linear(int n, int K, int N, const float *A, int lda, const float *W, int trans_b, const float *bias, float *Y, int ldy,
                   cudaStream_t st);

extern "C" int cb_linear_forward(int n, int ci, int co, const float *X, const float *W, const float *b, float *Y, void *stream)
{
    CB_REQUIRE(n >= 0 && ci > 0 && co > 0 && X && W && Y, CB_EINVAL, "cb_linear_forward: bad arguments");
    if (n == 0) return CB_OK;
    if (cb_tc_enabled() && cb_umma_shape_ok(n, ci, co, X, Y, ci, co)) {      // tcgen05 + TMEM (umma_linear.cu)
        cb_umma_linear(n, ci, co, X, ci, W, 0, b, Y, co, (cudaStream_t)stream);
        CB_CUDA_CHECK("cb_linear_forward");
        return CB_OK;
    }
    if (cb_tc_enabled()) {
        cb_tc_linear_forward(n, ci, co, X, W, b, Y, (cudaStream_t)stream);
        CB_COUNT(1);
        CB_CUDA_CHECK("cb_linear_forward");
        return CB_OK;
    }
    launch_skinny(n, ci, co, X, W, ci, 1, b, Y, (cudaStream_t)stream);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_linear_forward");
    return CB_OK;
}

extern "C" int cb_linear_dgrad(int n, int ci, int co, const float *G, const float *W, float *dX, void *stream)
{
    CB_REQUIRE(n >= 0 && ci > 0 && co > 0 && G && W && dX, CB_EINVAL, "cb_linear_dgrad: bad arguments");
    if (n == 0) return CB_OK;
    if (cb_tc_enabled() && cb_umma_shape_ok(n, co, ci, G, dX, co, ci)) {     // dX = G (n x co) . W (co x ci)
        cb_umma_linear(n, co, ci, G, co, W, 1, nullptr, dX, ci, (cudaStream_t)stream);
        CB_CUDA_CHECK("cb_linear_dgrad");
        return CB_OK;
    }
    if (cb_tc_enabled()) {
        cb_tc_linear_dgrad(n, ci, co, G, W, dX, (cudaStream_t)stream);
        CB_COUNT(1);
        CB_CUDA_CHECK("cb_linear_dgrad");
        return CB_OK;
    }
    launch_skinny(n, co, ci, G, W, ci, 0, nullptr, dX, (cudaStream_t)stream);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_linear_dgrad");
    return CB_OK;
}

extern "C" int cb_linear_wgrad(int n, int ci, int co, const float *X, const float *G, float *dW, float *db, void *stream)
{
    CB_REQUIRE(n >= 0 && ci > 0 && co > 0 && X && G && dW, CB_EINVAL, "cb_linear_wgrad: bad arguments");
    CB_REQUIRE(cb_tc_enabled() || (ci * co <= 64 * LG_THREADS && co <= LG_THREADS), CB_EUNSUPPORTED,
               "cb_linear_wgrad: ci*co=%d too large for the SIMT kernel", ci * co);
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)ci * co, st);
    if (db) cudaMemsetAsync(db, 0, sizeof(float) * (size_t)co, st);
    if (n == 0) return CB_OK;
    if (cb_tc_enabled()) {
        cb_tc_linear_wgrad(n, ci, co, X, G, dW, db, nullptr, nullptr, st);
        CB_COUNT(3);
        CB_CUDA_CHECK("cb_linear_wgrad");
        return CB_OK;
    }
    int blocks = 148 * 4;
    int rpb = (n + blocks - 1) / blocks;
    rpb = (rpb + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
    blocks = (n + rpb - 1) / rpb;
    const size_t smem = (size_t)WG_ROWS * (ci + co) * sizeof(float);
    const int opt = (ci * co + LG_THREADS - 1) / LG_THREADS;
    if (ci % 4 == 0 && co % 4 == 0 && ((uintptr_t)X | (uintptr_t)G) % 16 == 0) {
        const int ppt = (ci * co / 16 + LG_THREADS - 1) / LG_THREADS;
#define WG4_LAUNCH(P)                                                                                           \
    do {                                                                                                        \
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_skinny_wgrad4<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        k_skinny_wgrad4<P><<<blocks, LG_THREADS, smem, st>>>(n, ci, co, X, G, dW, db, rpb);                     \
    } while (0)
        if (ppt <= 1) WG4_LAUNCH(1);
        else if (ppt <= 2) WG4_LAUNCH(2);
        else WG4_LAUNCH(4);
#undef WG4_LAUNCH
        CB_COUNT(3);
        CB_CUDA_CHECK("cb_linear_wgrad");
        return CB_OK;
    }
#define WG_LAUNCH(O)                                                                                           \
    do {                                                                                                       \
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_skinny_wgrad<O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        k_skinny_wgrad<O><<<blocks, LG_THREADS, smem, st>>>(n, ci, co, X, G, dW, db, rpb);                     \
    } while (0)
    if (opt <= 1) WG_LAUNCH(1);
    else if (opt <= 4) WG_LAUNCH(4);
    else if (opt <= 8) WG_LAUNCH(8);
    else if (opt <= 16) WG_LAUNCH(16);
    else if (opt <= 32) WG_LAUNCH(32);
    else WG_LAUNCH(64);
#undef WG_LAUNCH
    CB_COUNT(3);
    CB_CUDA_CHECK("cb_linear_wgrad");
    return CB_OK;
}
