// tc_gemm.cu — tall-skinny FP32 linear layers on the tensor cores with 3xTF32 error compensation.
// Same operators as linear_ops.cu (reference: the nn.Linear calls of pytorch/model/blocks.py:33,72,76,108,127-131
// and the heads' MLPs), used for the shapes that dominate the step (n >= 8192 rows, 6..512 channels).
//
// Why tensor cores for an HBM-bound problem: the FP32 SIMT kernels of linear_ops.cu need ~3 issue slots per FMA
// (LDS + FFMA + address math) and run 4-6x above the HBM time of these shapes; on the tensor pipe the math is
// free and the kernels become memory-bound.  Why 3xTF32: parity with the FP32 reference is 1e-4 on features and
// loss, a plain TF32 product (10-bit mantissa) does not hold it through 40 layers.  Every operand x is split into
// hi = tf32(x) and lo = tf32(x - hi); a*b ~= hi*hi + lo*hi + hi*lo with FP32 accumulation (the lo*lo term is
// below FP32 rounding), which keeps the product error at ~2^-21 relative.
//   forward : Y[n,co]  = X[n,ci] W[co,ci]^T + b          (A row-major, B = W as stored: "col" operand)
//   dgrad   : dX[n,ci] = G[n,co] W[co,ci]                (B row-major [K][N], transposed while staging)
//   wgrad   : dW[co,ci] = G^T X, db[co] = sum_rows G     (reduction over rows: split over blocks, float atomics)
// Instruction: mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 (operands staged in shared memory by the
// block; no tcgen05/TMEM here on purpose: tiles are 128 x {32,64} x 32 and the kernels are bandwidth-bound).
#include "common.cuh"
#include "tf32.cuh"

#define TG_BM 128
#define TG_BK 32
#define TG_LDS 36          // padded k-stride of the staged tiles: bank = (4*row + k) mod 32 -> conflict-free fragments
#define TG_THREADS 256

// C[n x N] = A[n x K] * B (+ bias).  TRANS_B: B[k][j] = W[j*ldw + k] (forward), else W[k*ldw + j] (dgrad).
// Block: 128 rows x BN columns (BN = 32 | 64), 8 warps x 16 rows; blockIdx.y tiles N.  Persistent over row tiles:
// the A chunk of the NEXT (tile, k-chunk) is prefetched into registers while the current one is multiplied, and
// the (hi, lo)-split B tile stays resident in shared memory when K <= 32 (the common case: one k-chunk).
// EPI = 1 (backward of relu(bn(.)) fused into the epilogue, used by the PointTransformer layer): with z (n x N) the
// pre-BatchNorm activation and bnp = [scale | shift | mean | invstd] (N each), C = acc * [z*scale + shift > 0] and the
// per-column sums of C and of C * xhat (xhat = (z - mean) * invstd) are accumulated into sums[2][N] (double).
#define TG_STAGES 3
__device__ __forceinline__ void tg_cp_async16(void *smem, const void *gmem, bool valid)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;                   // 0 source bytes -> the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem), "r"(bytes) : "memory");
}
template <int BN> struct TgSmem {
    static constexpr size_t bytes = (size_t)TG_STAGES * TG_BM * TG_LDS * 4 + (size_t)2 * BN * TG_LDS * 4;
};

template <int BN, bool TRANS_B, int EPI>
__global__ void __launch_bounds__(TG_THREADS, 3) k_tc_gemm(int n, int K, int N, const float *__restrict__ A,
                                                           const float *__restrict__ W, int ldw,
                                                           const float *__restrict__ bias, float *__restrict__ C,
                                                           const float *__restrict__ z, const float *__restrict__ bnp,
                                                           double *__restrict__ sums)
{
    constexpr int NTILE = BN / 8;
    constexpr int APT = (TG_BM * TG_BK / 4) / TG_THREADS;     // 16-byte pieces of an A chunk per thread (4)
    extern __shared__ __align__(16) unsigned char tg_sm[];
    float (*As)[TG_BM][TG_LDS] = reinterpret_cast<float (*)[TG_BM][TG_LDS]>(tg_sm);                               // [STAGES]
    unsigned (*Bh)[TG_LDS] = reinterpret_cast<unsigned (*)[TG_LDS]>(tg_sm + (size_t)TG_STAGES * TG_BM * TG_LDS * 4);   // [BN]
    unsigned (*Bl)[TG_LDS] = Bh + BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int col0 = blockIdx.y * BN;
    const bool vecA = (K % 4 == 0) && (((uintptr_t)A & 15) == 0);
    const int nchunks = (K + TG_BK - 1) / TG_BK;
    const long long ntiles = ((long long)n + TG_BM - 1) / TG_BM;
    const long long my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long total_it = my_tiles * nchunks;

    // stage the A chunk of iteration `it` (row tile it / nchunks of this CTA, k-chunk it % nchunks) into slot it % STAGES
    auto issueA = [&](long long it) {
        const long long tile = blockIdx.x + (it / nchunks) * gridDim.x;
        const int k0 = (int)(it % nchunks) * TG_BK;
        const int slot = (int)(it % TG_STAGES);
        if (vecA) {
#pragma unroll
            for (int i = 0; i < APT; i++) {
                const int e = tid + i * TG_THREADS;
                const int r = e >> 3, kq = (e & 7) * 4;
                const long long row = tile * TG_BM + r;
                const bool ok = row < n && k0 + kq < K;
                tg_cp_async16(&As[slot][r][kq], ok ? (const void *)(A + row * K + k0 + kq) : (const void *)A, ok);
            }
        } else {
            for (int e = tid; e < TG_BM * TG_BK; e += TG_THREADS) {
                const int r = e >> 5, kk = e & 31;
                const long long row = tile * TG_BM + r;
                As[slot][r][kk] = (row < n && k0 + kk < K) ? __ldg(A + row * K + k0 + kk) : 0.f;
            }
        }
    };
    auto stageB = [&](int chunk) {
        const int k0 = chunk * TG_BK;
        for (int e = tid; e < BN * TG_BK; e += TG_THREADS) {
            int j, kk;
            if (TRANS_B) { j = e >> 5; kk = e & 31; }
            else { kk = e / BN; j = e % BN; }                    // coalesced along j in global memory
            float v = 0.f;
            if (col0 + j < N && k0 + kk < K)
                v = TRANS_B ? __ldg(W + (size_t)(col0 + j) * ldw + k0 + kk) : __ldg(W + (size_t)(k0 + kk) * ldw + col0 + j);
            unsigned hi, lo;
            tg_split(v, hi, lo);
            Bh[j][kk] = hi; Bl[j][kk] = lo;
        }
    };

    // EPI == 1: per-column BatchNorm constants and running column sums of this thread's output columns
    float esc[EPI ? NTILE : 1][2], esh[EPI ? NTILE : 1][2], emu[EPI ? NTILE : 1][2], eiv[EPI ? NTILE : 1][2];
    float sa[EPI ? NTILE : 1][2], sb[EPI ? NTILE : 1][2];
    if (EPI == 1) {
#pragma unroll
        for (int j = 0; j < NTILE; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int col = col0 + j * 8 + 2 * t + e;
                const bool ok = col < N;
                esc[j][e] = ok ? __ldg(bnp + col) : 0.f; esh[j][e] = ok ? __ldg(bnp + N + col) : 0.f;
                emu[j][e] = ok ? __ldg(bnp + 2 * N + col) : 0.f; eiv[j][e] = ok ? __ldg(bnp + 3 * N + col) : 0.f;
                sa[j][e] = 0.f; sb[j][e] = 0.f;
            }
    }
    if (nchunks == 1) stageB(0);                 // resident for the whole kernel (made visible by the first barrier below)
    for (int c = 0; c < TG_STAGES - 1; c++) {
        if (c < total_it) issueA(c);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    float acc[NTILE][4];
    float2 zz[EPI ? NTILE : 1][2];
    for (long long it = 0; it < total_it; it++) {
        const long long tile = blockIdx.x + (it / nchunks) * gridDim.x;
        const int chunk = (int)(it % nchunks);
        const long long row0 = tile * TG_BM;
        if (chunk == 0) {
#pragma unroll
            for (int j = 0; j < NTILE; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
            // EPI == 1: the pre-activation of this thread's outputs, fetched now so that its HBM latency hides under the MMAs
            if (EPI == 1) {
#pragma unroll
                for (int j = 0; j < NTILE; j++)
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const long long row = row0 + warp * 16 + g + 8 * h;
                        const int col = col0 + j * 8 + 2 * t;
                        zz[j][h] = make_float2(0.f, 0.f);
                        if (row < n && col < N) zz[j][h] = __ldg(reinterpret_cast<const float2 *>(z + row * N + col));
                    }
            }
        }
        asm volatile("cp.async.wait_group %0;" ::"n"(TG_STAGES - 2) : "memory");     // this iteration's A chunk has landed
        __syncthreads();                      // ... for every thread; the slot of iteration it-1 (and the B tile) are free again
        if (it + TG_STAGES - 1 < total_it) issueA(it + TG_STAGES - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (nchunks > 1) {
            stageB(chunk);
            __syncthreads();
        }
        const int slot = (int)(it % TG_STAGES);
#pragma unroll
        for (int ks = 0; ks < TG_BK; ks += 8) {
            if (chunk * TG_BK + ks >= K) break;          // zero padding beyond K (uniform)
            unsigned ah[4], al[4];
            tg_split(As[slot][warp * 16 + g][ks + t], ah[0], al[0]);
            tg_split(As[slot][warp * 16 + g + 8][ks + t], ah[1], al[1]);
            tg_split(As[slot][warp * 16 + g][ks + t + 4], ah[2], al[2]);
            tg_split(As[slot][warp * 16 + g + 8][ks + t + 4], ah[3], al[3]);
#pragma unroll
            for (int j = 0; j < NTILE; j++) {
                const unsigned bh0 = Bh[j * 8 + g][ks + t], bh1 = Bh[j * 8 + g][ks + t + 4];
                const unsigned bl0 = Bl[j * 8 + g][ks + t], bl1 = Bl[j * 8 + g][ks + t + 4];
                tg_mma(acc[j], al, bh0, bh1);      // small terms first
                tg_mma(acc[j], ah, bl0, bl1);
                tg_mma(acc[j], ah, bh0, bh1);
            }
        }
        if (chunk != nchunks - 1) continue;
        // ---- epilogue: c0,c1 -> (row g, cols 2t,2t+1); c2,c3 -> row g+8
#pragma unroll
        for (int j = 0; j < NTILE; j++) {
            const int col = col0 + j * 8 + 2 * t;
            float b0 = 0.f, b1 = 0.f;
            if (bias) { if (col < N) b0 = __ldg(bias + col); if (col + 1 < N) b1 = __ldg(bias + col + 1); }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const long long row = row0 + warp * 16 + g + 8 * h;
                if (row >= n) continue;
                float *dst = C + row * N + col;
                float v0 = acc[j][2 * h] + b0, v1 = acc[j][2 * h + 1] + b1;
                if (EPI == 1) {
                    // N is even and col is even: (col, col+1) are both valid or both out of range
                    if (col < N) {
                        const float2 zv = zz[j][h];
                        v0 = (zv.x * esc[j][0] + esh[j][0] > 0.f) ? v0 : 0.f;
                        v1 = (zv.y * esc[j][1] + esh[j][1] > 0.f) ? v1 : 0.f;
                        sa[j][0] += v0; sa[j][1] += v1;
                        sb[j][0] += v0 * ((zv.x - emu[j][0]) * eiv[j][0]);
                        sb[j][1] += v1 * ((zv.y - emu[j][1]) * eiv[j][1]);
                    }
                }
                if (col + 1 < N && ((N & 1) == 0)) *reinterpret_cast<float2 *>(dst) = make_float2(v0, v1);
                else { if (col < N) dst[0] = v0; if (col + 1 < N) dst[1] = v1; }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (EPI == 1) {
        // column sums: reduce over the 8 row-lanes (g) and the 8 warps, then one double atomic per column and block
        __shared__ float ecomb[2][BN];
        for (int i = tid; i < 2 * BN; i += TG_THREADS) (&ecomb[0][0])[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NTILE; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                float a = sa[j][e], b = sb[j][e];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    a += __shfl_xor_sync(CB_FULL_MASK, a, o);
                    b += __shfl_xor_sync(CB_FULL_MASK, b, o);
                }
                if (g == 0) { atomicAdd(&ecomb[0][j * 8 + 2 * t + e], a); atomicAdd(&ecomb[1][j * 8 + 2 * t + e], b); }
            }
        __syncthreads();
        for (int i = tid; i < BN; i += TG_THREADS)
            if (col0 + i < N) {
                atomicAdd(sums + col0 + i, (double)ecomb[0][i]);
                atomicAdd(sums + N + col0 + i, (double)ecomb[1][i]);
            }
    }
}

// wgrad: dW[co][ci] += sum_{rows of this block} G[row][co] * X[row][ci]   (A = G^T: M = co, B = X: N = ci, K = rows)
// Block tile 64 (co) x 64 (ci); warps 4 (co) x 2 (ci): each 16 x 32.  blockIdx.y / z tile co / ci, blockIdx.x = row chunk.
#define TW_T 64
#define TW_LDS 72          // (72 mod 32) = 8: bank = (8*k + m) mod 32 -> conflict-free fragments
// xsc / xsh (optional, ci each): the X operand is relu(X * xsc + xsh) — the post-BatchNorm activation recomputed from the
// stored pre-activation when the fragment is read (PointTransformer layer: dW3 = dw2^T relu(bn2(w0))).
// The kernel is a pure stream over the rows (0.1-0.7 GFLOP, 40-340 MB): 32-row chunks of G and X are moved by a
// TW_STAGES-deep cp.async pipeline (raw FP32; the TF32 hi / lo split happens in registers when a fragment is read), so
// every CTA keeps TW_STAGES - 1 chunks of loads in flight and 3 CTAs per SM cover the HBM latency.
#define TW_STAGES 4
__device__ __forceinline__ void tw_cp_async16(void *smem, const void *gmem, bool valid)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;                   // 0 source bytes -> the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(TG_THREADS, 3) k_tc_wgrad(int n, int ci, int co, const float *__restrict__ X,
                                                            const float *__restrict__ G, float *__restrict__ dW,
                                                            float *__restrict__ db, int rows_per_block,
                                                            const float *__restrict__ xsc, const float *__restrict__ xsh)
{
    extern __shared__ __align__(16) float tw_sm[];
    float (*Gs)[TG_BK][TW_LDS] = reinterpret_cast<float (*)[TG_BK][TW_LDS]>(tw_sm);                              // [STAGES]
    float (*Xs)[TG_BK][TW_LDS] = reinterpret_cast<float (*)[TG_BK][TW_LDS]>(tw_sm + TW_STAGES * TG_BK * TW_LDS);   // [STAGES]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    const int m0 = blockIdx.y * TW_T, n0 = blockIdx.z * TW_T;
    const bool active = (m0 + wm * 16 < co) && (n0 + wn * 32 < ci);     // warps whose 16 x 32 patch is pure padding idle
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    float accb = 0.f;
    const long long r_begin = (long long)blockIdx.x * rows_per_block;
    long long r_end = r_begin + rows_per_block;
    if (r_end > n) r_end = n;
    const int nchunks = r_end > r_begin ? (int)((r_end - r_begin + TG_BK - 1) / TG_BK) : 0;
    const bool vec = (ci % 4 == 0) && (co % 4 == 0) && ((((uintptr_t)X | (uintptr_t)G) & 15) == 0);
    // relu(bn(.)) prologue constants of this thread's B-fragment columns
    float psc[4], psh[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int c = n0 + wn * 32 + j * 8 + g;
        psc[j] = (xsc && c < ci) ? __ldg(xsc + c) : 1.f;
        psh[j] = (xsc && c < ci) ? __ldg(xsh + c) : 0.f;
    }
    auto issue = [&](int chunk) {                       // stage chunk -> slot chunk % TW_STAGES (zero beyond the edges)
        const int slot = chunk % TW_STAGES;
        const long long r0 = r_begin + (long long)chunk * TG_BK;
        if (vec) {
#pragma unroll
            for (int i = 0; i < (TG_BK * (TW_T / 4)) / TG_THREADS; i++) {
                const int e = tid + i * TG_THREADS;
                const int r = e >> 4, q = (e & 15) * 4;
                const long long row = r0 + r;
                const bool okg = row < r_end && m0 + q < co, okx = row < r_end && n0 + q < ci;
                tw_cp_async16(&Gs[slot][r][q], okg ? (const void *)(G + row * co + m0 + q) : (const void *)G, okg);
                tw_cp_async16(&Xs[slot][r][q], okx ? (const void *)(X + row * ci + n0 + q) : (const void *)X, okx);
            }
        } else {
            for (int e = tid; e < TG_BK * TW_T; e += TG_THREADS) {
                const int r = e >> 6, q = e & 63;
                const long long row = r0 + r;
                Gs[slot][r][q] = (row < r_end && m0 + q < co) ? __ldg(G + row * co + m0 + q) : 0.f;
                Xs[slot][r][q] = (row < r_end && n0 + q < ci) ? __ldg(X + row * ci + n0 + q) : 0.f;
            }
        }
    };
    // prologue: TW_STAGES - 1 chunks in flight (one commit group per chunk, empty groups keep the count uniform)
    for (int c = 0; c < TW_STAGES - 1; c++) {
        if (c < nchunks) issue(c);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int chunk = 0; chunk < nchunks; chunk++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(TW_STAGES - 2) : "memory");     // chunk's group has landed
        __syncthreads();                                 // ... for every thread, and slot (chunk-1) % STAGES is free again
        if (chunk + TW_STAGES - 1 < nchunks) issue(chunk + TW_STAGES - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const int slot = chunk % TW_STAGES;
        if (active) {
#pragma unroll
            for (int ks = 0; ks < TG_BK; ks += 8) {
                unsigned ah[4], al[4];
                tg_split(Gs[slot][ks + t][wm * 16 + g], ah[0], al[0]);
                tg_split(Gs[slot][ks + t][wm * 16 + g + 8], ah[1], al[1]);
                tg_split(Gs[slot][ks + t + 4][wm * 16 + g], ah[2], al[2]);
                tg_split(Gs[slot][ks + t + 4][wm * 16 + g + 8], ah[3], al[3]);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c = wn * 32 + j * 8 + g;
                    float x0 = Xs[slot][ks + t][c], x1 = Xs[slot][ks + t + 4][c];
                    if (xsc) { x0 = fmaxf(x0 * psc[j] + psh[j], 0.f); x1 = fmaxf(x1 * psc[j] + psh[j], 0.f); }
                    unsigned bh0, bl0, bh1, bl1;
                    tg_split(x0, bh0, bl0);
                    tg_split(x1, bh1, bl1);
                    tg_mma(acc[j], al, bh0, bh1);
                    tg_mma(acc[j], ah, bl0, bl1);
                    tg_mma(acc[j], ah, bh0, bh1);
                }
            }
        }
        if (db && blockIdx.z == 0 && tid < TW_T) {
            float s2 = 0.f;
#pragma unroll 8
            for (int r = 0; r < TG_BK; r++) s2 += Gs[slot][r][tid];
            accb += s2;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (active) {
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int m = m0 + wm * 16 + g + 8 * h, c = n0 + wn * 32 + j * 8 + 2 * t;
                if (m < co) {
                    if (c < ci) atomicAdd(dW + (size_t)m * ci + c, acc[j][2 * h]);
                    if (c + 1 < ci) atomicAdd(dW + (size_t)m * ci + c + 1, acc[j][2 * h + 1]);
                }
            }
    }
    if (db && blockIdx.z == 0 && tid < TW_T && m0 + tid < co) atomicAdd(db + m0 + tid, accb);
}

static int g_tc_enabled = 1;
extern "C" int cb_linear_set_tensor_cores(int on) { g_tc_enabled = on ? 1 : 0; return g_tc_enabled; }
int cb_tc_enabled() { return g_tc_enabled; }

template <bool TRANS_B, int EPI>
static void tc_launch(int n, int K, int N, const float *A, const float *W, int ldw, const float *bias, float *C, const float *z,
                      const float *bnp, double *sums, cudaStream_t st)
{
    const int ntiles = (n + TG_BM - 1) / TG_BM;
    // column tiles of 64 (or 32 when that wastes fewer padded columns); the A tile of further column tiles comes from L2
    const int w64 = (N + 63) / 64 * 64 - N, w32 = (N + 31) / 32 * 32 - N;
    // (the fused-epilogue variant keeps per-column constants and sums in registers: 32-wide tiles only; its A operand is
    //  the narrow (rows x c/8) matrix, so re-reading it per column tile is cheap)
    const bool use32 = N <= 32 || w32 < w64 || EPI == 1;
    const int gy = use32 ? (N + 31) / 32 : (N + 63) / 64;
    int gx = (148 * 2 + gy - 1) / gy;                        // persistent: ~2 CTAs per SM in total
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    int gx3 = (148 * 3 + gy - 1) / gy;                       // persistent: ~3 CTAs per SM in total
    if (gx3 > ntiles) gx3 = ntiles;
    if (gx3 < 1) gx3 = 1;
    if (use32) {
        static bool set32 = false;
        if (!set32) { cudaFuncSetAttribute(k_tc_gemm<32, TRANS_B, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TgSmem<32>::bytes); set32 = true; }
        k_tc_gemm<32, TRANS_B, EPI><<<dim3(gx3, gy), TG_THREADS, TgSmem<32>::bytes, st>>>(n, K, N, A, W, ldw, bias, C, z, bnp, sums);
    } else {
        static bool set64 = false;
        if (!set64) { cudaFuncSetAttribute(k_tc_gemm<64, TRANS_B, EPI == 1 ? 0 : EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TgSmem<64>::bytes); set64 = true; }
        k_tc_gemm<64, TRANS_B, EPI == 1 ? 0 : EPI><<<dim3(gx3, gy), TG_THREADS, TgSmem<64>::bytes, st>>>(n, K, N, A, W, ldw, bias, C, z, bnp, sums);
    }
}

void cb_tc_linear_forward(int n, int ci, int co, const float *X, const float *W, const float *b, float *Y, cudaStream_t st)
{
    tc_launch<true, 0>(n, ci, co, X, W, ci, b, Y, nullptr, nullptr, nullptr, st);
}
void cb_tc_linear_dgrad(int n, int ci, int co, const float *G, const float *W, float *dX, cudaStream_t st)
{
    tc_launch<false, 0>(n, co, ci, G, W, ci, nullptr, dX, nullptr, nullptr, nullptr, st);
}
// dX = (G W) * [z * scale + shift > 0] with the BatchNorm-backward column sums (see k_tc_gemm, EPI = 1); ci must be even
void cb_tc_linear_dgrad_relu_bn(int n, int ci, int co, const float *G, const float *W, float *dX, const float *z,
                                const float *bnp, double *sums, cudaStream_t st)
{
    tc_launch<false, 1>(n, co, ci, G, W, ci, nullptr, dX, z, bnp, sums, st);
}
// dW and db must be zero-filled by the caller; xsc / xsh: optional relu(bn(.)) prologue on X
void cb_tc_linear_wgrad(int n, int ci, int co, const float *X, const float *G, float *dW, float *db, const float *xsc,
                        const float *xsh, cudaStream_t st)
{
    const int ty = (co + TW_T - 1) / TW_T, tz = (ci + TW_T - 1) / TW_T;
    int blocks = (148 * 3) / (ty * tz);                       // 3 CTAs per SM: enough loads in flight for the narrow operands
    if (blocks < 8) blocks = 8;
    int rpb = (n + blocks - 1) / blocks;
    rpb = (rpb + TG_BK - 1) / TG_BK * TG_BK;
    blocks = (n + rpb - 1) / rpb;
    const size_t smem = (size_t)2 * TW_STAGES * TG_BK * TW_LDS * sizeof(float);     // 73.7 KB: three CTAs per SM
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_tc_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    k_tc_wgrad<<<dim3(blocks, ty, tz), TG_THREADS, smem, st>>>(n, ci, co, X, G, dW, db, rpb, xsc, xsh);
}
