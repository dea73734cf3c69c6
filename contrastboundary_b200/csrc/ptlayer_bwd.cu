// ptlayer_bwd.cu — a4: fused PointTransformer local aggregation, BACKWARD.
// Forward definitions: ptlayer_fwd.cu (reference blocks.py:31-44).  With G = d out (n,c):
//   dval = G a                         da[k,j]  = sum_{c%CS=j} G[c] (x_v[idx]+pr)[k,c]        (B1)
//   dw4  = a (da - sum_k a da)         dW4 += dw4 (x) v, db4 ; dv = W4^T dw4 ; dy3 = dv [y3>0]   (B2)
//   dw2  = BN3'(dy3)                   dW3 += dw2 (x) u ; du = W3^T dw2 ; dy2 = du [y2>0]        (B4)
//   dw0  = BN2'(dy2)                   dx_k[idx] += dw0 ; dx_q = -sum_k dw0 ; dx_v[idx] += dval
//   dpr  = dw0 + dval                  dW2 += dpr (x) g1, db2 ; dg1 = W2^T dpr ; dy1 = dg1 [g1>0] (B5)
//   dh1  = BN1'(dy1)                   dW1 += dh1 (x) r, db1                                     (B6)
// BN'(dy) = gamma*invstd*(dy - mean(dy) - xhat*mean(dy*xhat)) needs two global sums per BatchNorm, hence
// the kernel boundaries.  (n,k,c) quantities are recomputed from the L2-resident tables, never stored;
// only (n,k,c/8) and (n,k,3) tensors touch HBM.  Scatter-adds use vector float atomics (red.v4).
#include "ptlayer.cuh"

#define PT_THREADS 256
#define PT_WARPS (PT_THREADS / 32)
#define PT_KMAX 32

struct PtSmall {
    float w1[9], b1[3], sc1[3], sh1[3];
};
__device__ __forceinline__ PtSmall pt_small_load(const float *__restrict__ d)
{
    PtSmall s;
#pragma unroll
    for (int i = 0; i < 9; i++) s.w1[i] = __ldg(d + i);
#pragma unroll
    for (int i = 0; i < 3; i++) { s.b1[i] = __ldg(d + 9 + i); s.sc1[i] = __ldg(d + 12 + i); s.sh1[i] = __ldg(d + 15 + i); }
    return s;
}
__device__ __forceinline__ void pt_g1h(const PtSmall &sp, float rx, float ry, float rz, float (&g)[3], float (&h)[3])
{
#pragma unroll
    for (int a = 0; a < 3; a++) {
        h[a] = sp.w1[3 * a] * rx + sp.w1[3 * a + 1] * ry + sp.w1[3 * a + 2] * rz + sp.b1[a];
        g[a] = fmaxf(h[a] * sp.sc1[a] + sp.sh1[a], 0.f);
    }
}
int pt_grid(int n);

// coefficient block of one BatchNorm backward: [kk | ma | mb] (C each): dw = kk*(dy - ma - xhat*mb)
__global__ void k_pt_bn_coef(const double *__restrict__ sums /* [2][C]: sum dy, sum dy*xhat */, double count, int C,
                             const float *__restrict__ gamma, const float *__restrict__ invstd, int training,
                             float *__restrict__ coef, float *__restrict__ dgamma, float *__restrict__ dbeta)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double sa = sums[c], sb = sums[C + c];
    coef[c] = gamma[c] * invstd[c];
    coef[C + c] = training ? (float)(sa / count) : 0.f;
    coef[2 * C + c] = training ? (float)(sb / count) : 0.f;
    dgamma[c] = (float)sb;
    dbeta[c] = (float)sa;
}

// ---------------------------------------------------------------------------------------------
// B1: da[n,k,j] = sum_{c % CS = j} G[n,c] * (x_v[idx] + pr)[n,k,c]
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(PT_THREADS) k_pt_bwd_da(int n, int k, int ld, const float *__restrict__ rel,
                                                          const int *__restrict__ idx, const float *__restrict__ xv,
                                                          const float *__restrict__ w2p, const float *__restrict__ b2p,
                                                          const float *__restrict__ smalld, const float *__restrict__ G,
                                                          float *__restrict__ D)
{
    const PtSmall sp = pt_small_load(smalld);
    using M = PtMap<C>;
    constexpr int VW = M::VW, NS = M::NS, CS = M::CS, L = M::L;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float wa[NS][VW], wb[NS][VW], wc[NS][VW], bb[NS][VW];
#pragma unroll
    for (int s = 0; s < NS; s++)
#pragma unroll
        for (int v = 0; v < VW; v++) {
            const int ch = M::ch(lane, s, v);
            wa[s][v] = w2p[3 * ch]; wb[s][v] = w2p[3 * ch + 1]; wc[s][v] = w2p[3 * ch + 2]; bb[s][v] = b2p[ch];
        }
    const int warps = gridDim.x * PT_WARPS;
    for (int pt = blockIdx.x * PT_WARPS + wib; pt < n; pt += warps) {
        float g[NS][VW];
#pragma unroll
        for (int s = 0; s < NS; s++) pt_load<VW>(G + (size_t)pt * C + M::ch(lane, s, 0), g[s]);
        for (int kk = 0; kk < k; kk++) {
            const size_t row = (size_t)pt * k + kk;
            const int j = __ldg(idx + row);
            float g1[3], h1[3];
            pt_g1h(sp, __ldg(rel + 3 * row), __ldg(rel + 3 * row + 1), __ldg(rel + 3 * row + 2), g1, h1);
            float part[VW];
#pragma unroll
            for (int v = 0; v < VW; v++) part[v] = 0.f;
#pragma unroll
            for (int s = 0; s < NS; s++) {
                float x[VW];
                pt_load<VW>(xv + (size_t)j * ld + M::ch(lane, s, 0), x);
#pragma unroll
                for (int v = 0; v < VW; v++) {
                    const float pr = wa[s][v] * g1[0] + wb[s][v] * g1[1] + wc[s][v] * g1[2] + bb[s][v];
                    part[v] += g[s][v] * (x[v] + pr);
                }
            }
#pragma unroll
            for (int off = L; off < 32; off <<= 1)
#pragma unroll
                for (int v = 0; v < VW; v++) part[v] += __shfl_xor_sync(CB_FULL_MASK, part[v], off);
            if (lane < L) pt_store<VW>(D + row * CS + lane * VW, part);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// B2 (cs-space): softmax backward, linear_w[5] backward, relu/bn3 partial sums.  D: da -> dy3 in place.
// A block walks tiles of PB points; thread (pt, j) / (pt, i).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT_THREADS) k_pt_bwd_softmax(int n, int k, int CS, const float *__restrict__ w2,
                                                               const float *__restrict__ a, const float *__restrict__ bn3,
                                                               const float *__restrict__ w4, float *__restrict__ D,
                                                               float *__restrict__ gW4, float *__restrict__ gb4,
                                                               double *__restrict__ sums3)
{
    extern __shared__ float smb[];
    const int PB = PT_THREADS / CS;                 // points per tile
    float *sw4 = smb;                               // [CS][CS+1]
    float *sdw = sw4 + CS * (CS + 1);               // [PB][k][CS]   dw4
    float *sv = sdw + PB * k * CS;                  // [PB][k][CS]   v = relu(bn3(w2))
    float *sacc = sv + PB * k * CS;                 // [2*CS] block partial sums of dy3, dy3*xhat ; [CS] db4
    for (int i = threadIdx.x; i < CS * CS; i += PT_THREADS) sw4[(i / CS) * (CS + 1) + i % CS] = w4[i];
    for (int i = threadIdx.x; i < 3 * CS; i += PT_THREADS) sacc[i] = 0.f;
    const int tp = threadIdx.x / CS, tj = threadIdx.x % CS;
    const float sc3 = bn3[tj], sh3 = bn3[CS + tj], mean3 = bn3[2 * CS + tj], inv3 = bn3[3 * CS + tj];
    // dW4 accumulators: pairs (j,i) = e / CS, e % CS for e = tid, tid + 256, ...
    float accw[16];
#pragma unroll
    for (int e = 0; e < 16; e++) accw[e] = 0.f;
    float acc_a = 0.f, acc_b = 0.f, acc_db4 = 0.f;
    __syncthreads();
    const int tiles = (n + PB - 1) / PB;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int pt = tile * PB + tp;
        const bool act = pt < n;
        // phase A: thread (pt, j): softmax backward
        if (act) {
            float dot = 0.f;
            for (int kk = 0; kk < k; kk++) {
                const size_t e = ((size_t)pt * k + kk) * CS + tj;
                dot += __ldg(a + e) * D[e];
            }
            for (int kk = 0; kk < k; kk++) {
                const size_t e = ((size_t)pt * k + kk) * CS + tj;
                const float d = __ldg(a + e) * (D[e] - dot);
                sdw[(tp * k + kk) * CS + tj] = d;
                acc_db4 += d;
                sv[(tp * k + kk) * CS + tj] = fmaxf(__ldg(w2 + e) * sc3 + sh3, 0.f);
            }
        } else {
            for (int kk = 0; kk < k; kk++) { sdw[(tp * k + kk) * CS + tj] = 0.f; sv[(tp * k + kk) * CS + tj] = 0.f; }
        }
        __syncthreads();
        // phase B: thread (pt, i = tj): dv = W4^T dw4, dy3 = dv [v > 0]
        if (act) {
            for (int kk = 0; kk < k; kk++) {
                const float *dwr = sdw + (tp * k + kk) * CS;
                float dv = 0.f;
                for (int j = 0; j < CS; j++) dv += sw4[j * (CS + 1) + tj] * dwr[j];
                const size_t e = ((size_t)pt * k + kk) * CS + tj;
                const float dy = sv[(tp * k + kk) * CS + tj] > 0.f ? dv : 0.f;
                D[e] = dy;
                acc_a += dy;
                acc_b += dy * ((__ldg(w2 + e) - mean3) * inv3);
            }
        }
        // phase C: dW4[j][i] += sum_rows dw4[row][j] * v[row][i]
        const int rows = PB * k;
        const int npairs = CS * CS;
        if (npairs <= PT_THREADS) {
            // few pairs (CS <= 16): every thread takes one pair and a strided subset of the tile's rows
            const int pair = threadIdx.x % npairs, sub = threadIdx.x / npairs, nsub = PT_THREADS / npairs;
            const int j = pair / CS, i = pair % CS;
            float s = 0.f;
            for (int r = sub; r < rows; r += nsub) s += sdw[r * CS + j] * sv[r * CS + i];
            accw[0] += s;
        } else {
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const int pair = threadIdx.x + e * PT_THREADS;
                if (pair < npairs) {
                    const int j = pair / CS, i = pair % CS;
                    float s = 0.f;
                    for (int r = 0; r < rows; r++) s += sdw[r * CS + j] * sv[r * CS + i];
                    accw[e] += s;
                }
            }
        }
        __syncthreads();
    }
    atomicAdd(&sacc[tj], acc_a);
    atomicAdd(&sacc[CS + tj], acc_b);
    atomicAdd(&sacc[2 * CS + tj], acc_db4);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * CS; i += PT_THREADS) atomicAdd(sums3 + i, (double)sacc[i]);
    for (int i = threadIdx.x; i < CS; i += PT_THREADS) atomicAdd(gb4 + i, sacc[2 * CS + i]);
    if (CS * CS <= PT_THREADS) {
        atomicAdd(gW4 + threadIdx.x % (CS * CS), accw[0]);
    } else {
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int pair = threadIdx.x + e * PT_THREADS;
            if (pair < CS * CS) atomicAdd(gW4 + pair, accw[e]);
        }
    }
}

// dw2 of one row from dy3 (D), w2 and the bn3 coefficients
__device__ __forceinline__ float pt_dw2(float dy3, float w2v, float mean3, float inv3, float kk3, float ma3, float mb3)
{
    return kk3 * (dy3 - ma3 - (w2v - mean3) * inv3 * mb3);
}

// ---------------------------------------------------------------------------------------------
// B4 (tensor-core path): dw2 = BN3'(dy3) materialised as a (rows, CS) matrix — the narrow operand of the two
// tensor-core GEMMs dW3 = dw2^T relu(bn2(w0)) and dy2 = (dw2 W3) [y2 > 0] (tc_gemm.cu) — and db3 = sum_rows dw2.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT_THREADS) k_pt_dw2(long long rows, int CS, const float *__restrict__ D,
                                                       const float *__restrict__ w2buf, const float *__restrict__ bn3,
                                                       const float *__restrict__ coef3, float *__restrict__ dw2,
                                                       float *__restrict__ gb3)
{
    // thread -> column i = tid % CS (CS divides 256), rows strided
    const int i = threadIdx.x % CS, rsub = threadIdx.x / CS, RPB = PT_THREADS / CS;
    const float m3 = bn3[2 * CS + i], i3 = bn3[3 * CS + i], k3 = coef3[i], a3 = coef3[CS + i], b3 = coef3[2 * CS + i];
    float acc = 0.f;
    for (long long r = (long long)blockIdx.x * RPB + rsub; r < rows; r += (long long)gridDim.x * RPB) {
        const float v = pt_dw2(__ldg(D + r * CS + i), __ldg(w2buf + r * CS + i), m3, i3, k3, a3, b3);
        dw2[r * CS + i] = v;
        acc += v;
    }
    __shared__ float comb[64];
    if (threadIdx.x < 64) comb[threadIdx.x] = 0.f;
    __syncthreads();
    atomicAdd(&comb[i], acc);
    __syncthreads();
    if (threadIdx.x < CS) atomicAdd(gb3 + threadIdx.x, comb[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------
// B4: dW3[i][c] += sum_rows dw2[row,i] * u[row,c] ; sums of dy2 = (W3^T dw2) [y2>0] and dy2*xhat2.
// Thread (row-in-group, channel): CT = min(C,256) channels per block column, RG = 256/CT rows at once.
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(PT_THREADS) k_pt_bwd_dw3(int n, int k, int ld, const float *__restrict__ rel,
                                                           const int *__restrict__ idx, const float *__restrict__ xq,
                                                           const float *__restrict__ xk, const float *__restrict__ w2p,
                                                           const float *__restrict__ b2p, const float *__restrict__ smalld,
                                                           const float *__restrict__ bn2, const float *__restrict__ bn3,
                                                           const float *__restrict__ coef3, const float *__restrict__ w3,
                                                           const float *__restrict__ w2buf, const float *__restrict__ D,
                                                           float *__restrict__ gW3, float *__restrict__ gb3,
                                                           double *__restrict__ sums2)
{
    constexpr int CS = C / 8;
    constexpr int CT = C < 256 ? C : 256;          // channels per block column
    constexpr int RG = PT_THREADS / CT;            // row slots per block
    constexpr int TR = 64;                         // rows staged per iteration (TR / RG rows per thread)
    const PtSmall sp = pt_small_load(smalld);
    __shared__ __align__(16) float sdw2[TR][CS];
    __shared__ float sg1[TR][3];
    __shared__ int sidx[TR];
    const int tr = threadIdx.x / CT;
    const int ch = blockIdx.y * CT + threadIdx.x % CT;
    float w3c[CS], acc[CS];
#pragma unroll
    for (int i = 0; i < CS; i++) { w3c[i] = w3[i * C + ch]; acc[i] = 0.f; }
    const float wa = w2p[3 * ch], wb = w2p[3 * ch + 1], wc = w2p[3 * ch + 2], bb = b2p[ch];
    const float sc2 = bn2[ch], sh2 = bn2[C + ch], mean2 = bn2[2 * C + ch], inv2 = bn2[3 * C + ch];
    float sa = 0.f, sb = 0.f, sdb3 = 0.f;
    const long long rows = (long long)n * k;
    const long long tiles = (rows + TR - 1) / TR;
    for (long long tI = blockIdx.x; tI < tiles; tI += gridDim.x) {
        const long long row0 = tI * TR;
        for (int e = threadIdx.x; e < TR * CS; e += PT_THREADS) {
            const int r = e / CS, i = e % CS;
            const long long row = row0 + r;
            float v = 0.f;
            if (row < rows)
                v = pt_dw2(__ldg(D + row * CS + i), __ldg(w2buf + row * CS + i), bn3[2 * CS + i], bn3[3 * CS + i], coef3[i],
                           coef3[CS + i], coef3[2 * CS + i]);
            sdw2[r][i] = v;
        }
        if (threadIdx.x < TR) {
            const long long row = row0 + threadIdx.x;
            float g1[3] = {0, 0, 0}, h1[3];
            int j = 0;
            if (row < rows) {
                pt_g1h(sp, __ldg(rel + 3 * row), __ldg(rel + 3 * row + 1), __ldg(rel + 3 * row + 2), g1, h1);
                j = __ldg(idx + row);
            }
            sg1[threadIdx.x][0] = g1[0]; sg1[threadIdx.x][1] = g1[1]; sg1[threadIdx.x][2] = g1[2];
            sidx[threadIdx.x] = j;
        }
        __syncthreads();
        // gather phase first (UB independent loads in flight), then the math
        constexpr int RPT = TR / RG;                   // rows per thread and tile
        constexpr int UB = RPT < 8 ? RPT : 8;
#pragma unroll 1
        for (int u0 = 0; u0 < RPT; u0 += UB) {
            float xkv[UB], xqv[UB];
#pragma unroll
            for (int u = 0; u < UB; u++) {
                const int r = tr + (u0 + u) * RG;
                const long long row = row0 + r;
                xkv[u] = 0.f; xqv[u] = 0.f;
                if (row < rows) {
                    xkv[u] = __ldg(xk + (size_t)sidx[r] * ld + ch);
                    xqv[u] = __ldg(xq + (size_t)(row / k) * ld + ch);
                }
            }
#pragma unroll
            for (int u = 0; u < UB; u++) {
                const int r = tr + (u0 + u) * RG;
                if (row0 + r < rows) {
                    const float pr = wa * sg1[r][0] + wb * sg1[r][1] + wc * sg1[r][2] + bb;
                    const float w0 = xkv[u] - xqv[u] + pr;
                    const float y2 = w0 * sc2 + sh2;
                    const float uu = fmaxf(y2, 0.f);
                    float du = 0.f;
#pragma unroll
                    for (int i = 0; i < CS; i += 4) {
                        const float4 d4 = *reinterpret_cast<const float4 *>(&sdw2[r][i]);   // one LDS.128 feeds 8 FMAs
                        du += w3c[i] * d4.x + w3c[i + 1] * d4.y + w3c[i + 2] * d4.z + w3c[i + 3] * d4.w;
                        acc[i] += d4.x * uu; acc[i + 1] += d4.y * uu; acc[i + 2] += d4.z * uu; acc[i + 3] += d4.w * uu;
                    }
                    const float dy2 = y2 > 0.f ? du : 0.f;
                    sa += dy2;
                    sb += dy2 * ((w0 - mean2) * inv2);
                }
            }
        }
        if (blockIdx.y == 0 && threadIdx.x < CS) {       // db3 = sum dw2
            float t = 0.f;
            for (int r = 0; r < TR; r++) t += sdw2[r][threadIdx.x];
            sdb3 += t;
        }
        __syncthreads();
    }
    // combine the block's row slots through shared memory (RG > 1), then global atomics
    const int lc = threadIdx.x % CT;
    if constexpr (RG > 1) {
        __shared__ float comb[CS * CT + 2 * CT];
        for (int i = threadIdx.x; i < CS * CT + 2 * CT; i += PT_THREADS) comb[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < CS; i++) atomicAdd(&comb[i * CT + lc], acc[i]);
        atomicAdd(&comb[CS * CT + lc], sa);
        atomicAdd(&comb[CS * CT + CT + lc], sb);
        __syncthreads();
        for (int e = threadIdx.x; e < CS * CT; e += PT_THREADS) {
            const int i = e / CT, c2 = e % CT;
            atomicAdd(gW3 + (size_t)i * C + blockIdx.y * CT + c2, comb[e]);
        }
        for (int e = threadIdx.x; e < CT; e += PT_THREADS) {
            atomicAdd(sums2 + blockIdx.y * CT + e, (double)comb[CS * CT + e]);
            atomicAdd(sums2 + C + blockIdx.y * CT + e, (double)comb[CS * CT + CT + e]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < CS; i++) atomicAdd(gW3 + (size_t)i * C + ch, acc[i]);
        atomicAdd(sums2 + ch, (double)sa);
        atomicAdd(sums2 + C + ch, (double)sb);
        (void)lc;
    }
    if (blockIdx.y == 0 && threadIdx.x < CS) atomicAdd(gb3 + threadIdx.x, sdb3);
}

// ---------------------------------------------------------------------------------------------
// B5: main c-space backward (warp per point)
// ---------------------------------------------------------------------------------------------
// STORED: the pre-BatchNorm activation w0 (n,k,C) was kept by the forward pass and dy2 = (W3^T dw2) [y2 > 0] (n,k,C) was
// produced by the tensor-core GEMM (tc_gemm.cu, fused relu-bn epilogue): the kernel streams both instead of gathering
// x_k and re-doing the c x c/8 contraction (w0s = w0, dy2s = dy2; w3 / w2buf / D / coef3 / bn3 are then unused).
template <int C, bool STORED>
__global__ void __launch_bounds__(PT_THREADS) k_pt_bwd_main(int n, int k, int ld, const float *__restrict__ rel,
                                                            const int *__restrict__ idx, const float *__restrict__ xq,
                                                            const float *__restrict__ xk, const float *__restrict__ w2p,
                                                            const float *__restrict__ b2p, const float *__restrict__ smalld,
                                                            const float *__restrict__ bn1ms /* mean1[3] invstd1[3] */,
                                                            const float *__restrict__ bn2, const float *__restrict__ bn3,
                                                            const float *__restrict__ coef2, const float *__restrict__ coef3,
                                                            const float *__restrict__ w3, const float *__restrict__ w2buf,
                                                            const float *__restrict__ abuf, const float *__restrict__ D,
                                                            const float *__restrict__ G, float *__restrict__ gxq,
                                                            float *__restrict__ gxk, float *__restrict__ gxv,
                                                            float *__restrict__ gW2, float *__restrict__ gb2,
                                                            float *__restrict__ dy1, double *__restrict__ sums1,
                                                            const float *__restrict__ w0s, const float *__restrict__ dy2s)
{
    const PtSmall sp = pt_small_load(smalld);
    using M = PtMap<C>;
    constexpr int VW = M::VW, NS = M::NS, CS = M::CS;
    constexpr int JPL = CS > 32 ? CS / 32 : 1;          // dw2 entries per lane
    extern __shared__ __align__(16) float sm_w3[];     // [CS][C]
    if (!STORED)
        for (int i = threadIdx.x; i < CS * C; i += PT_THREADS) sm_w3[i] = w3[i];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float wa[NS][VW], wb[NS][VW], wc[NS][VW], bb[NS][VW], sc[NS][VW], sh[NS][VW], mu[NS][VW], iv[NS][VW];
    float k2[NS][VW], ma2[NS][VW], mb2[NS][VW];
    float aW2[NS][VW][3], ab2[NS][VW];
#pragma unroll
    for (int s = 0; s < NS; s++)
#pragma unroll
        for (int v = 0; v < VW; v++) {
            const int ch = M::ch(lane, s, v);
            wa[s][v] = w2p[3 * ch]; wb[s][v] = w2p[3 * ch + 1]; wc[s][v] = w2p[3 * ch + 2]; bb[s][v] = b2p[ch];
            sc[s][v] = bn2[ch]; sh[s][v] = bn2[C + ch]; mu[s][v] = bn2[2 * C + ch]; iv[s][v] = bn2[3 * C + ch];
            k2[s][v] = coef2[ch]; ma2[s][v] = coef2[C + ch]; mb2[s][v] = coef2[2 * C + ch];
            aW2[s][v][0] = aW2[s][v][1] = aW2[s][v][2] = 0.f; ab2[s][v] = 0.f;
        }
    // bn3 pieces for the dw2 entries this lane owns: i = lane + 32*t
    float m3[JPL], i3[JPL], k3[JPL], a3[JPL], b3c[JPL];
#pragma unroll
    for (int t = 0; t < JPL; t++) {
        const int i = (lane + 32 * t) % CS;
        m3[t] = 0.f; i3[t] = 0.f; k3[t] = 0.f; a3[t] = 0.f; b3c[t] = 0.f;
        if (!STORED) { m3[t] = bn3[2 * CS + i]; i3[t] = bn3[3 * CS + i]; k3[t] = coef3[i]; a3[t] = coef3[CS + i]; b3c[t] = coef3[2 * CS + i]; }
    }
    const float mean1[3] = {bn1ms[0], bn1ms[1], bn1ms[2]}, inv1[3] = {bn1ms[3], bn1ms[4], bn1ms[5]};
    float s1a[3] = {0, 0, 0}, s1b[3] = {0, 0, 0};
    __syncthreads();
    const int warps = gridDim.x * PT_WARPS;
    for (int pt = blockIdx.x * PT_WARPS + wib; pt < n; pt += warps) {
        float q[NS][VW], g[NS][VW], dq[NS][VW];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            if (!STORED) pt_load<VW>(xq + (size_t)pt * ld + M::ch(lane, s, 0), q[s]);
            pt_load<VW>(G + (size_t)pt * C + M::ch(lane, s, 0), g[s]);
#pragma unroll
            for (int v = 0; v < VW; v++) dq[s][v] = 0.f;
        }
        // STORED: the streamed operands of row kk+1 are fetched (HBM latency) while row kk is processed
        float w0n[STORED ? NS : 1][VW], dun[STORED ? NS : 1][VW];
        if (STORED) {
#pragma unroll
            for (int s = 0; s < NS; s++) {
                pt_load<VW>(w0s + (size_t)pt * k * C + M::ch(lane, s, 0), w0n[s]);
                pt_load<VW>(dy2s + (size_t)pt * k * C + M::ch(lane, s, 0), dun[s]);
            }
        }
        for (int kk = 0; kk < k; kk++) {
            const size_t row = (size_t)pt * k + kk;
            const int j = __ldg(idx + row);
            float g1[3], h1[3];
            pt_g1h(sp, __ldg(rel + 3 * row), __ldg(rel + 3 * row + 1), __ldg(rel + 3 * row + 2), g1, h1);
            float w0c[STORED ? NS : 1][VW], duc[STORED ? NS : 1][VW];
            if (STORED) {
#pragma unroll
                for (int s = 0; s < NS; s++) {
#pragma unroll
                    for (int v = 0; v < VW; v++) { w0c[s][v] = w0n[s][v]; duc[s][v] = dun[s][v]; }
                    if (kk + 1 < k) {
                        pt_load<VW>(w0s + (row + 1) * C + M::ch(lane, s, 0), w0n[s]);
                        pt_load<VW>(dy2s + (row + 1) * C + M::ch(lane, s, 0), dun[s]);
                    }
                }
            }
            // dw2 entries owned by this lane
            float dw2o[JPL];
#pragma unroll
            for (int t = 0; t < JPL; t++) {
                const int i = (lane + 32 * t) % CS;
                dw2o[t] = 0.f;
                if (!STORED) dw2o[t] = pt_dw2(__ldg(D + row * CS + i), __ldg(w2buf + row * CS + i), m3[t], i3[t], k3[t], a3[t], b3c[t]);
            }
            float dg[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int s = 0; s < NS; s++) {
                float x[VW], aw[VW], w0[VW], du[VW];
                pt_load<VW>(abuf + row * CS + (M::ch(lane, s, 0) % CS), aw);
                if (STORED) {
#pragma unroll
                    for (int v = 0; v < VW; v++) { w0[v] = w0c[s][v]; du[v] = duc[s][v]; }   // du already masked by [y2 > 0]
                } else {
                    pt_load<VW>(xk + (size_t)j * ld + M::ch(lane, s, 0), x);
#pragma unroll
                    for (int v = 0; v < VW; v++) {
                        const float pr = wa[s][v] * g1[0] + wb[s][v] * g1[1] + wc[s][v] * g1[2] + bb[s][v];
                        w0[v] = x[v] - q[s][v] + pr;
                        du[v] = 0.f;
                    }
                }
#pragma unroll
                for (int i = 0; i < (STORED ? 0 : CS); i++) {
                    const float d = __shfl_sync(CB_FULL_MASK, dw2o[i / 32], i % 32);
                    const float *wp = sm_w3 + i * C + M::ch(lane, s, 0);
                    float wv[VW];
                    if (VW == 4) { const float4 t = *reinterpret_cast<const float4 *>(wp); wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w; }
                    else if (VW == 2) { const float2 t = *reinterpret_cast<const float2 *>(wp); wv[0] = t.x; wv[1] = t.y; }
                    else wv[0] = wp[0];
#pragma unroll
                    for (int v = 0; v < VW; v++) du[v] += wv[v] * d;
                }
                float dw0[VW], dval[VW];
#pragma unroll
                for (int v = 0; v < VW; v++) {
                    const float y2 = w0[v] * sc[s][v] + sh[s][v];
                    const float dy2 = (STORED || y2 > 0.f) ? du[v] : 0.f;
                    dw0[v] = k2[s][v] * (dy2 - ma2[s][v] - (w0[v] - mu[s][v]) * iv[s][v] * mb2[s][v]);
                    dval[v] = g[s][v] * aw[v];
                    dq[s][v] -= dw0[v];
                    const float dpr = dw0[v] + dval[v];
                    aW2[s][v][0] += dpr * g1[0]; aW2[s][v][1] += dpr * g1[1]; aW2[s][v][2] += dpr * g1[2];
                    ab2[s][v] += dpr;
                    dg[0] += dpr * wa[s][v]; dg[1] += dpr * wb[s][v]; dg[2] += dpr * wc[s][v];
                }
                pt_red_add<VW>(gxk + (size_t)j * ld + M::ch(lane, s, 0), dw0);
                float xvadd[VW];
#pragma unroll
                for (int v = 0; v < VW; v++) xvadd[v] = dval[v];
                pt_red_add<VW>(gxv + (size_t)j * ld + M::ch(lane, s, 0), xvadd);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                dg[0] += __shfl_xor_sync(CB_FULL_MASK, dg[0], o);
                dg[1] += __shfl_xor_sync(CB_FULL_MASK, dg[1], o);
                dg[2] += __shfl_xor_sync(CB_FULL_MASK, dg[2], o);
            }
            if (lane == 0) {
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const float d = g1[a] > 0.f ? dg[a] : 0.f;
                    dy1[3 * row + a] = d;
                    s1a[a] += d;
                    s1b[a] += d * ((h1[a] - mean1[a]) * inv1[a]);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < NS; s++) pt_store<VW>(gxq + (size_t)pt * ld + M::ch(lane, s, 0), dq[s]);
    }
    // block combine of dW2 / db2 / bn1 sums
    __shared__ float comb[4 * C + 6];
    for (int i = threadIdx.x; i < 4 * C + 6; i += PT_THREADS) comb[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NS; s++)
#pragma unroll
        for (int v = 0; v < VW; v++) {
            const int ch = M::ch(lane, s, v);
            atomicAdd(&comb[3 * ch], aW2[s][v][0]); atomicAdd(&comb[3 * ch + 1], aW2[s][v][1]);
            atomicAdd(&comb[3 * ch + 2], aW2[s][v][2]); atomicAdd(&comb[3 * C + ch], ab2[s][v]);
        }
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) { atomicAdd(&comb[4 * C + a], s1a[a]); atomicAdd(&comb[4 * C + 3 + a], s1b[a]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * C; i += PT_THREADS) atomicAdd(gW2 + i, comb[i]);
    for (int i = threadIdx.x; i < C; i += PT_THREADS) atomicAdd(gb2 + i, comb[3 * C + i]);
    if (threadIdx.x < 6) atomicAdd(sums1 + threadIdx.x, (double)comb[4 * C + threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------
// B6: linear_p[0] backward through bn1 (thread per row)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT_THREADS) k_pt_bwd_bn1(long long rows, const float *__restrict__ rel,
                                                           const float *__restrict__ smalld,
                                                           const float *__restrict__ bn1ms, const float *__restrict__ coef1,
                                                           const float *__restrict__ dy1, float *__restrict__ gW1,
                                                           float *__restrict__ gb1)
{
    const PtSmall sp = pt_small_load(smalld);
    float acc[12];
#pragma unroll
    for (int i = 0; i < 12; i++) acc[i] = 0.f;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
        const float rx = __ldg(rel + 3 * r), ry = __ldg(rel + 3 * r + 1), rz = __ldg(rel + 3 * r + 2);
        float g1[3], h1[3];
        pt_g1h(sp, rx, ry, rz, g1, h1);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float xh = (h1[a] - bn1ms[a]) * bn1ms[3 + a];
            const float dh = coef1[a] * (__ldg(dy1 + 3 * r + a) - coef1[3 + a] - xh * coef1[6 + a]);
            acc[3 * a] += dh * rx; acc[3 * a + 1] += dh * ry; acc[3 * a + 2] += dh * rz;
            acc[9 + a] += dh;
        }
    }
    __shared__ float red[12][PT_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CB_FULL_MASK, v, o);
        if (lane == 0) red[i][w] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float t = 0.f;
        for (int i = 0; i < PT_WARPS; i++) t += red[threadIdx.x][i];
        atomicAdd(threadIdx.x < 9 ? gW1 + threadIdx.x : gb1 + (threadIdx.x - 9), t);
    }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
extern "C" size_t cb_pt_bwd_scratch_floats(int n, int k, int c)
{
    const size_t cs = (size_t)c / 8;
    const size_t dbl = 2 * (2 * cs + 2 * (size_t)c + 6 + 8);          // double region, in floats
    const size_t coef = 3 * cs + 3 * (size_t)c + 9 + 16;
    // + dw2 (n,k,cs) and dy2 (n,k,c) of the tensor-core path
    return dbl + coef + (size_t)n * k * cs + (size_t)n * k * 3 + 64 + (size_t)n * k * cs + (size_t)n * k * c + 64;
}

int cb_pt_mma_enabled();
void cb_tc_linear_dgrad_relu_bn(int n, int ci, int co, const float *G, const float *W, float *dX, const float *z,
                                const float *bnp, double *sums, cudaStream_t st);
void cb_tc_linear_wgrad(int n, int ci, int co, const float *X, const float *G, float *dW, float *db, const float *xsc,
                        const float *xsh, cudaStream_t st);

template <int C>
static int pt_backward_c(int n, int k, int ld, const CbPtLayer *L, const float *rel, const int *idx, const float *xq,
                         const float *xk, const float *xv, const float *w2buf, const float *abuf, const float *bnbuf,
                         const float *G, float *gxq, float *gxk, float *gxv, float *gbuf, float *scratch, const float *w0buf,
                         cudaStream_t st)
{
    constexpr int CS = C / 8;
    const float *small = bnbuf, *bn1ms = bnbuf + 18, *bn2 = bnbuf + 24, *bn3 = bnbuf + 24 + 4 * C;
    // scratch carve-up
    double *sums3 = (double *)scratch;                 // [2][CS]
    double *sums2 = sums3 + 2 * CS;                    // [2][C]
    double *sums1 = sums2 + 2 * C;                     // [2][3]
    float *coef3 = (float *)(sums1 + 6 + 2);           // [3][CS]
    float *coef2 = coef3 + 3 * CS;                     // [3][C]
    float *coef1 = coef2 + 3 * C;                      // [3][3]
    float *D = coef1 + 9 + 7;                          // (n,k,CS)
    float *dy1 = D + (size_t)n * k * CS;               // (n,k,3)
    // gbuf layout
    float *gW1 = gbuf, *gb1 = gW1 + 9, *gg1 = gb1 + 3, *gbe1 = gg1 + 3, *gW2 = gbe1 + 3, *gb2 = gW2 + 3 * C,
          *gg2 = gb2 + C, *gbe2 = gg2 + C, *gW3 = gbe2 + C, *gb3 = gW3 + CS * C, *gg3 = gb3 + CS, *gbe3 = gg3 + CS,
          *gW4 = gbe3 + CS, *gb4 = gW4 + CS * CS;
    const double rows = (double)n * (double)k;
    cudaMemsetAsync(scratch, 0, sizeof(double) * (2 * CS + 2 * C + 8), st);
    const int grid = pt_grid(n);
    k_pt_bwd_da<C><<<grid, PT_THREADS, 0, st>>>(n, k, ld, rel, idx, xv, L->w2, L->b2, small, G, D);
    {
        const int PB = PT_THREADS / CS;
        const size_t smem = (size_t)(CS * (CS + 1) + 2 * PB * k * CS + 3 * CS) * sizeof(float);
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_pt_bwd_softmax, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int tiles = (n + PB - 1) / PB;
        int g2 = tiles < 148 * 2 ? (tiles < 1 ? 1 : tiles) : 148 * 2;
        k_pt_bwd_softmax<<<g2, PT_THREADS, smem, st>>>(n, k, CS, w2buf, abuf, bn3, L->w4, D, gW4, gb4, sums3);
    }
    k_pt_bn_coef<<<1, 128, 0, st>>>(sums3, rows, CS, L->bn3_weight, bn3 + 3 * CS, L->training, coef3, gg3, gbe3);
    const long long nrows = (long long)n * k;
    const bool stored = w0buf != nullptr && cb_pt_mma_enabled() && nrows < (1LL << 31);
    float *DW2 = dy1 + (((size_t)n * k * 3 + 3) & ~(size_t)3) + 16;   // (n,k,CS)   tensor-core path only (16-byte aligned)
    float *DY2 = DW2 + (size_t)n * k * CS + 16;                      // (n,k,C)
    if (stored) {
        // tensor cores (3xTF32): dw2 -> dW3 = dw2^T relu(bn2(w0)) ; dy2 = (dw2 W3) [y2 > 0] with the bn2-backward sums
        int gd = (int)((nrows * CS + PT_THREADS - 1) / PT_THREADS);
        if (gd > 148 * 8) gd = 148 * 8;
        if (gd < 1) gd = 1;
        k_pt_dw2<<<gd, PT_THREADS, 0, st>>>(nrows, CS, D, w2buf, bn3, coef3, DW2, gb3);
        cb_tc_linear_wgrad((int)nrows, C, CS, w0buf, DW2, gW3, nullptr, bn2, bn2 + C, st);
        cb_tc_linear_dgrad_relu_bn((int)nrows, C, CS, DW2, L->w3, DY2, w0buf, bn2, sums2, st);
    } else {
        constexpr int CT = C < 256 ? C : 256;
        const long long groups = ((long long)n * k + 63) / 64;
        int gx = (int)(groups < 148 * 4 ? (groups < 1 ? 1 : groups) : 148 * 4);
        dim3 g4(gx, C / CT);
        k_pt_bwd_dw3<C><<<g4, PT_THREADS, 0, st>>>(n, k, ld, rel, idx, xq, xk, L->w2, L->b2, small, bn2, bn3, coef3, L->w3, w2buf, D,
                                                   gW3, gb3, sums2);
    }
    k_pt_bn_coef<<<(C + 127) / 128, 128, 0, st>>>(sums2, rows, C, L->bn2_weight, bn2 + 3 * C, L->training, coef2, gg2, gbe2);
    if (stored) {
        k_pt_bwd_main<C, true><<<grid, PT_THREADS, 0, st>>>(n, k, ld, rel, idx, xq, xk, L->w2, L->b2, small, bn1ms, bn2, bn3, coef2,
                                                            coef3, L->w3, w2buf, abuf, D, G, gxq, gxk, gxv, gW2, gb2, dy1,
                                                            sums1, w0buf, DY2);
    } else {
        const size_t smem = (size_t)CS * C * sizeof(float);
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_pt_bwd_main<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_pt_bwd_main<C, false><<<grid, PT_THREADS, smem, st>>>(n, k, ld, rel, idx, xq, xk, L->w2, L->b2, small, bn1ms, bn2, bn3, coef2,
                                                                coef3, L->w3, w2buf, abuf, D, G, gxq, gxk, gxv, gW2, gb2, dy1,
                                                                sums1, nullptr, nullptr);
    }
    k_pt_bn_coef<<<1, 32, 0, st>>>(sums1, rows, 3, L->bn1_weight, bn1ms + 3, L->training, coef1, gg1, gbe1);
    {
        const long long r = (long long)n * k;
        int g6 = (int)((r + PT_THREADS - 1) / PT_THREADS);
        if (g6 > 148 * 4) g6 = 148 * 4;
        if (g6 < 1) g6 = 1;
        k_pt_bwd_bn1<<<g6, PT_THREADS, 0, st>>>(r, rel, small, bn1ms, coef1, dy1, gW1, gb1);
    }
    CB_COUNT(9);
    CB_CUDA_CHECK("cb_pt_layer_backward");
    return CB_OK;
}

extern "C" int cb_pt_layer_backward(int n, int k, int c, int ld, const CbPtLayer *L, const float *rel, const int *idx,
                                    const float *xq, const float *xk, const float *xv, const float *w2buf,
                                    const float *abuf, const float *bnbuf, const float *grad_out, float *grad_xq,
                                    float *grad_xk, float *grad_xv, float *grad_params, float *scratch, const float *w0buf,
                                    void *stream)
{
    CB_REQUIRE(n >= 0 && k >= 1 && k <= PT_KMAX, CB_EINVAL, "cb_pt_layer_backward: n=%d k=%d", n, k);
    CB_REQUIRE(L && rel && idx && xq && xk && xv && w2buf && abuf && bnbuf && grad_out && grad_xq && grad_xk && grad_xv &&
                   grad_params && scratch, CB_EINVAL, "cb_pt_layer_backward: NULL pointer");
    CB_REQUIRE(((uintptr_t)scratch & 15) == 0, CB_EINVAL, "cb_pt_layer_backward: scratch not 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return CB_OK;
    switch (c) {
#define PT_CASE(CC) case CC: return pt_backward_c<CC>(n, k, ld, L, rel, idx, xq, xk, xv, w2buf, abuf, bnbuf, grad_out, grad_xq, grad_xk, grad_xv, grad_params, scratch, w0buf, st);
        PT_CASE(32) PT_CASE(64) PT_CASE(128) PT_CASE(256) PT_CASE(512)
#undef PT_CASE
    default:
        cb_set_error("cb_pt_layer_backward: c=%d unsupported", c);
        return CB_EUNSUPPORTED;
    }
}
