// ptlayer_fwd.cu — a4: fused PointTransformer local aggregation, FORWARD.
// Reference math: PointTransformerLayer.forward, pytorch/model/blocks.py:31-44 (vector self-attention
// over the K neighbours), i.e. per point n and neighbour k (idx[n,k]):
//     r   = p[idx] - p[n]                                   (n,k,3)    geometry only
//     g1  = relu(bn1(W1 r + b1))                             linear_p[0..2]
//     pr  = W2 g1 + b2                                       linear_p[3]           (n,k,c)
//     w0  = x_k[idx] - x_q[n] + pr                           blocks.py:39
//     w2  = W3 relu(bn2(w0)) + b3                            linear_w[0..2]        (n,k,c/8)
//     a   = softmax_k( W4 relu(bn3(w2)) + b4 )               linear_w[3..5], :41
//     out = sum_k (x_v[idx] + pr)[c] * a[k, c % (c/8)]       :43
// The reference materialises ~10 (n,k,c) tensors per layer through ~25 kernels; here the (n,k,c)
// quantities only ever exist in registers: three gather passes over L2-resident feature tables
// (training-mode BatchNorm needs the global statistics of w0 and w2 before they can be consumed),
// with only the (n,k,c/8) tensors w2 and a written to HBM.  All math is FP32 SIMT: the dense
// contractions are tiny (<= 0.7 GFLOP per layer) and must stay within 1e-4 of the FP32 reference.
#include "ptlayer.cuh"

#define PT_THREADS 256
#define PT_WARPS (PT_THREADS / 32)

struct PtSmall {        // layer-constant small parameters (3-channel MLP): [w1(9) b1(3) sc1(3) sh1(3)] on the device
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

__device__ __forceinline__ void pt_g1(const PtSmall &sp, float rx, float ry, float rz, float (&g)[3])
{
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float h = sp.w1[3 * a] * rx + sp.w1[3 * a + 1] * ry + sp.w1[3 * a + 2] * rz + sp.b1[a];
        g[a] = fmaxf(h * sp.sc1[a] + sp.sh1[a], 0.f);
    }
}

// ---------------------------------------------------------------------------------------------
// geometry: rel = p[idx] - p[n] and the 9 moments of rel (sum r_a, sum r_a r_b), once per level
// ---------------------------------------------------------------------------------------------
__global__ void k_pt_rel(const float *__restrict__ p, const int *__restrict__ idx, long long rows, int k,
                         float *__restrict__ rel, double *__restrict__ mom)
{
    float s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
        const long long n = r / k;
        const int j = __ldg(idx + r);
        const float x = __ldg(p + 3 * j) - __ldg(p + 3 * n), y = __ldg(p + 3 * j + 1) - __ldg(p + 3 * n + 1),
                    z = __ldg(p + 3 * j + 2) - __ldg(p + 3 * n + 2);
        rel[3 * r] = x; rel[3 * r + 1] = y; rel[3 * r + 2] = z;
        s[0] += x; s[1] += y; s[2] += z;
        s[3] += x * x; s[4] += x * y; s[5] += x * z; s[6] += y * y; s[7] += y * z; s[8] += z * z;
    }
    __shared__ float red[9][PT_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 9; a++) {
        float v = s[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CB_FULL_MASK, v, o);
        if (lane == 0) red[a][w] = v;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double t = 0;
        for (int i = 0; i < PT_WARPS; i++) t += red[threadIdx.x][i];
        atomicAdd(mom + threadIdx.x, t);
    }
}

// bn1 statistics of h = W1 r + b1 from the moments of r (exact: h is affine in r)
__global__ void k_pt_bn1_prep(const double *__restrict__ mom, double count, const float *__restrict__ w1,
                              const float *__restrict__ b1, const float *__restrict__ gamma,
                              const float *__restrict__ beta, float *running_mean, float *running_var, float momentum,
                              float eps, int training, float *__restrict__ out /* sc[3] sh[3] mean[3] invstd[3] */)
{
    const int a = threadIdx.x;
    if (a >= 3) return;
    float mean, var;
    if (training) {
        const double mr[3] = {mom[0] / count, mom[1] / count, mom[2] / count};
        const double e2[3][3] = {{mom[3] / count, mom[4] / count, mom[5] / count},
                                 {mom[4] / count, mom[6] / count, mom[7] / count},
                                 {mom[5] / count, mom[7] / count, mom[8] / count}};
        double m = b1[a], v = 0;
        for (int i = 0; i < 3; i++) m += (double)w1[3 * a + i] * mr[i];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) v += (double)w1[3 * a + i] * (double)w1[3 * a + j] * (e2[i][j] - mr[i] * mr[j]);
        if (v < 0) v = 0;
        mean = (float)m; var = (float)v;
        const double unbiased = count > 1 ? v * count / (count - 1) : v;
        running_mean[a] = (1.f - momentum) * running_mean[a] + momentum * mean;
        running_var[a] = (1.f - momentum) * running_var[a] + momentum * (float)unbiased;
    } else {
        mean = running_mean[a]; var = running_var[a];
    }
    const float invstd = 1.0f / sqrtf(var + eps);
    const float sc = gamma[a] * invstd;
    out[a] = sc; out[3 + a] = beta[a] - mean * sc; out[6 + a] = mean; out[9 + a] = invstd;
}

// generic finalize: sums -> scale/shift/mean/invstd (+ running stats), C channels
__global__ void k_bn_finalize(const double *__restrict__ stats /* [2][C] */, double count, int C,
                              const float *__restrict__ gamma, const float *__restrict__ beta, float *running_mean,
                              float *running_var, float momentum, float eps, int training,
                              float *__restrict__ out /* [4][C] */)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float mean, var;
    if (training) {
        const double m = stats[c] / count;
        double v = stats[C + c] / count - m * m;
        if (v < 0) v = 0;
        mean = (float)m; var = (float)v;
        const double unbiased = count > 1 ? v * count / (count - 1) : v;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    } else {
        mean = running_mean[c]; var = running_var[c];
    }
    const float invstd = 1.0f / sqrtf(var + eps);
    const float sc = gamma[c] * invstd;
    out[c] = sc; out[C + c] = beta[c] - mean * sc; out[2 * C + c] = mean; out[3 * C + c] = invstd;
}

// ---------------------------------------------------------------------------------------------
// F1: statistics of w0 = x_k[idx] - x_q[n] + pr  (per channel sum / sum of squares)
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(PT_THREADS) k_pt_w0_stats(int n, int k, int ld, const float *__restrict__ rel,
                                                            const int *__restrict__ idx, const float *__restrict__ xq,
                                                            const float *__restrict__ xk, const float *__restrict__ w2p,
                                                            const float *__restrict__ b2p, const float *__restrict__ smalld,
                                                            double *__restrict__ stats, float *__restrict__ w0out)
{
    const PtSmall sp = pt_small_load(smalld);
    using M = PtMap<C>;
    constexpr int VW = M::VW, NS = M::NS;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float wa[NS][VW], wb[NS][VW], wc[NS][VW], bb[NS][VW], s1[NS][VW], s2[NS][VW];
#pragma unroll
    for (int s = 0; s < NS; s++)
#pragma unroll
        for (int v = 0; v < VW; v++) {
            const int ch = M::ch(lane, s, v);
            wa[s][v] = w2p[3 * ch]; wb[s][v] = w2p[3 * ch + 1]; wc[s][v] = w2p[3 * ch + 2]; bb[s][v] = b2p[ch];
            s1[s][v] = 0.f; s2[s][v] = 0.f;
        }
    const int warps = gridDim.x * PT_WARPS;
    for (int pt = blockIdx.x * PT_WARPS + wib; pt < n; pt += warps) {
        float q[NS][VW];
#pragma unroll
        for (int s = 0; s < NS; s++) pt_load<VW>(xq + (size_t)pt * ld + M::ch(lane, s, 0), q[s]);
        for (int kk = 0; kk < k; kk++) {
            const size_t row = (size_t)pt * k + kk;
            const int j = __ldg(idx + row);
            float g[3];
            pt_g1(sp, __ldg(rel + 3 * row), __ldg(rel + 3 * row + 1), __ldg(rel + 3 * row + 2), g);
#pragma unroll
            for (int s = 0; s < NS; s++) {
                float x[VW], w0v[VW];
                pt_load<VW>(xk + (size_t)j * ld + M::ch(lane, s, 0), x);
#pragma unroll
                for (int v = 0; v < VW; v++) {
                    const float pr = wa[s][v] * g[0] + wb[s][v] * g[1] + wc[s][v] * g[2] + bb[s][v];
                    const float w0 = x[v] - q[s][v] + pr;
                    w0v[v] = w0;
                    s1[s][v] += w0; s2[s][v] += w0 * w0;
                }
                if (w0out) pt_store<VW>(w0out + row * C + M::ch(lane, s, 0), w0v);   // kept for the backward pass
            }
        }
    }
    // combine the block's warps, then one double atomic per channel and block
    __shared__ float comb[2 * C];
    for (int i = threadIdx.x; i < 2 * C; i += PT_THREADS) comb[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NS; s++)
#pragma unroll
        for (int v = 0; v < VW; v++) {
            atomicAdd(&comb[M::ch(lane, s, v)], s1[s][v]);
            atomicAdd(&comb[C + M::ch(lane, s, v)], s2[s][v]);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += PT_THREADS) atomicAdd(stats + i, (double)comb[i]);
}

// ---------------------------------------------------------------------------------------------
// F2: w2 = W3 relu(bn2(w0)) + b3  -> (n,k,CS) and its per-channel statistics
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(PT_THREADS) k_pt_w2(int n, int k, int ld, const float *__restrict__ rel,
                                                      const int *__restrict__ idx, const float *__restrict__ xq,
                                                      const float *__restrict__ xk, const float *__restrict__ w2p,
                                                      const float *__restrict__ b2p, const float *__restrict__ smalld,
                                                      const float *__restrict__ bn2 /* [4][C] */,
                                                      const float *__restrict__ w3 /* [CS][C] */,
                                                      const float *__restrict__ b3, float *__restrict__ w2out,
                                                      double *__restrict__ stats /* [2][CS] */)
{
    const PtSmall sp = pt_small_load(smalld);
    using M = PtMap<C>;
    constexpr int VW = M::VW, NS = M::NS, CS = M::CS;
    extern __shared__ __align__(16) float sm_w3[];   // [CS][C]
    for (int i = threadIdx.x; i < CS * C; i += PT_THREADS) sm_w3[i] = w3[i];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float wa[NS][VW], wb[NS][VW], wc[NS][VW], bb[NS][VW], sc[NS][VW], sh[NS][VW];
#pragma unroll
    for (int s = 0; s < NS; s++)
#pragma unroll
        for (int v = 0; v < VW; v++) {
            const int ch = M::ch(lane, s, v);
            wa[s][v] = w2p[3 * ch]; wb[s][v] = w2p[3 * ch + 1]; wc[s][v] = w2p[3 * ch + 2]; bb[s][v] = b2p[ch];
            sc[s][v] = bn2[ch]; sh[s][v] = bn2[C + ch];
        }
    // halving-reduction bookkeeping: which outputs this lane ends up holding
    constexpr int H = CS >= 32 ? 5 : (CS == 16 ? 4 : (CS == 8 ? 3 : 2));   // halving steps = min(5, log2 CS)
    constexpr int MF = CS >> H;                                            // outputs per lane at the end
    int jbase = 0;
    {
        int m = CS;
#pragma unroll
        for (int st = 0; st < H; st++) {
            const int off = 16 >> st;
            m >>= 1;
            if (lane & off) jbase += m;
        }
    }
    const bool writer = (lane & ((1 << (5 - H)) - 1)) == 0;
    float t1[MF], t2[MF], bias3[MF];
#pragma unroll
    for (int i = 0; i < MF; i++) { t1[i] = 0.f; t2[i] = 0.f; bias3[i] = b3[jbase + i]; }
    __syncthreads();
    const int warps = gridDim.x * PT_WARPS;
    for (int pt = blockIdx.x * PT_WARPS + wib; pt < n; pt += warps) {
        float q[NS][VW];
#pragma unroll
        for (int s = 0; s < NS; s++) pt_load<VW>(xq + (size_t)pt * ld + M::ch(lane, s, 0), q[s]);
        for (int kk = 0; kk < k; kk++) {
            const size_t row = (size_t)pt * k + kk;
            const int j = __ldg(idx + row);
            float g[3];
            pt_g1(sp, __ldg(rel + 3 * row), __ldg(rel + 3 * row + 1), __ldg(rel + 3 * row + 2), g);
            float acc[CS];
#pragma unroll
            for (int i = 0; i < CS; i++) acc[i] = 0.f;
#pragma unroll
            for (int s = 0; s < NS; s++) {
                float x[VW], u[VW];
                pt_load<VW>(xk + (size_t)j * ld + M::ch(lane, s, 0), x);
#pragma unroll
                for (int v = 0; v < VW; v++) {
                    const float pr = wa[s][v] * g[0] + wb[s][v] * g[1] + wc[s][v] * g[2] + bb[s][v];
                    u[v] = fmaxf((x[v] - q[s][v] + pr) * sc[s][v] + sh[s][v], 0.f);
                }
#pragma unroll
                for (int i = 0; i < CS; i++) {
                    float wv[VW];
                    const float *wp = sm_w3 + i * C + M::ch(lane, s, 0);
                    if (VW == 4) { const float4 t = *reinterpret_cast<const float4 *>(wp); wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w; }
                    else if (VW == 2) { const float2 t = *reinterpret_cast<const float2 *>(wp); wv[0] = t.x; wv[1] = t.y; }
                    else wv[0] = wp[0];
#pragma unroll
                    for (int v = 0; v < VW; v++) acc[i] += wv[v] * u[v];
                }
            }
            // cross-lane sum by recursive halving: after H steps each lane holds MF complete outputs
            {
                int m = CS;
#pragma unroll
                for (int st = 0; st < 5; st++) {
                    const int off = 16 >> st;
                    if (st < H) {
                        const int half = m >> 1;
                        const bool upper = (lane & off) != 0;
#pragma unroll
                        for (int i = 0; i < CS / 2; i++) {
                            if (i < half) {
                                const float send = upper ? acc[i] : acc[i + half];
                                const float keep = upper ? acc[i + half] : acc[i];
                                acc[i] = keep + __shfl_xor_sync(CB_FULL_MASK, send, off);
                            }
                        }
                        m = half;
                    } else {
#pragma unroll
                        for (int i = 0; i < MF; i++) acc[i] += __shfl_xor_sync(CB_FULL_MASK, acc[i], off);
                    }
                }
            }
            if (writer) {
#pragma unroll
                for (int i = 0; i < MF; i++) {
                    const float o = acc[i] + bias3[i];
                    w2out[row * CS + jbase + i] = o;
                    t1[i] += o; t2[i] += o * o;
                }
            }
        }
    }
    __shared__ float comb[2 * CS];
    for (int i = threadIdx.x; i < 2 * CS; i += PT_THREADS) comb[i] = 0.f;
    __syncthreads();
    if (writer) {
#pragma unroll
        for (int i = 0; i < MF; i++) {
            atomicAdd(&comb[jbase + i], t1[i]);
            atomicAdd(&comb[CS + jbase + i], t2[i]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * CS; i += PT_THREADS) atomicAdd(stats + i, (double)comb[i]);
}

// ---------------------------------------------------------------------------------------------
// F3: a = softmax_k( W4 relu(bn3(w2)) + b4 )   (cs-space; one thread per (point, j))
// ---------------------------------------------------------------------------------------------
#define PT_KMAX 32
__global__ void k_pt_softmax(int n, int k, int CS, const float *__restrict__ w2, const float *__restrict__ bn3,
                             const float *__restrict__ w4, const float *__restrict__ b4, float *__restrict__ a)
{
    extern __shared__ float sm4[];                 // w4 [CS][CS+1], sc3[CS], sh3[CS]
    float *sw = sm4, *ssc = sm4 + CS * (CS + 1), *ssh = ssc + CS;
    for (int i = threadIdx.x; i < CS * CS; i += blockDim.x) sw[(i / CS) * (CS + 1) + i % CS] = w4[i];
    for (int i = threadIdx.x; i < CS; i += blockDim.x) { ssc[i] = bn3[i]; ssh[i] = bn3[CS + i]; }
    __syncthreads();
    const long long total = (long long)n * CS;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long pt = e / CS;
        const int j = (int)(e - pt * CS);
        float w[PT_KMAX];
        float mx = -3.0e38f;
        const float bj = __ldg(b4 + j);
#pragma unroll 1
        for (int kk = 0; kk < k; kk++) {
            const float *row = w2 + ((size_t)pt * k + kk) * CS;
            float acc = bj;
            for (int i = 0; i < CS; i++) acc += sw[j * (CS + 1) + i] * fmaxf(__ldg(row + i) * ssc[i] + ssh[i], 0.f);
            w[kk] = acc;
            mx = fmaxf(mx, acc);
        }
        float sum = 0.f;
        for (int kk = 0; kk < k; kk++) { w[kk] = expf(w[kk] - mx); sum += w[kk]; }
        const float inv = 1.0f / sum;
        for (int kk = 0; kk < k; kk++) a[((size_t)pt * k + kk) * CS + j] = w[kk] * inv;
    }
}

// ---------------------------------------------------------------------------------------------
// F4: out[n,c] = sum_k (x_v[idx] + pr)[c] * a[n,k,c % CS]
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(PT_THREADS) k_pt_aggregate(int n, int k, int ld, const float *__restrict__ rel,
                                                             const int *__restrict__ idx, const float *__restrict__ xv,
                                                             const float *__restrict__ w2p, const float *__restrict__ b2p,
                                                             const float *__restrict__ smalld, const float *__restrict__ a,
                                                             float *__restrict__ out)
{
    const PtSmall sp = pt_small_load(smalld);
    using M = PtMap<C>;
    constexpr int VW = M::VW, NS = M::NS, CS = M::CS;
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
        float acc[NS][VW];
#pragma unroll
        for (int s = 0; s < NS; s++)
#pragma unroll
            for (int v = 0; v < VW; v++) acc[s][v] = 0.f;
        for (int kk = 0; kk < k; kk++) {
            const size_t row = (size_t)pt * k + kk;
            const int j = __ldg(idx + row);
            float g[3];
            pt_g1(sp, __ldg(rel + 3 * row), __ldg(rel + 3 * row + 1), __ldg(rel + 3 * row + 2), g);
#pragma unroll
            for (int s = 0; s < NS; s++) {
                float x[VW], aw[VW];
                pt_load<VW>(xv + (size_t)j * ld + M::ch(lane, s, 0), x);
                pt_load<VW>(a + row * CS + (M::ch(lane, s, 0) % CS), aw);
#pragma unroll
                for (int v = 0; v < VW; v++) {
                    const float pr = wa[s][v] * g[0] + wb[s][v] * g[1] + wc[s][v] * g[2] + bb[s][v];
                    acc[s][v] += (x[v] + pr) * aw[v];
                }
            }
        }
#pragma unroll
        for (int s = 0; s < NS; s++) pt_store<VW>(out + (size_t)pt * C + M::ch(lane, s, 0), acc[s]);
    }
}

// ---------------------------------------------------------------------------------------------
// host entry points
// ---------------------------------------------------------------------------------------------
__global__ void k_pt_pack_small(const float *__restrict__ w1, const float *__restrict__ b1, float *__restrict__ bnbuf)
{
    // bnbuf[0..12) = w1, b1 ; bnbuf[12..24) = sc1 sh1 mean1 invstd1 (written by k_pt_bn1_prep)
    const int t = threadIdx.x;
    if (t < 9) bnbuf[t] = w1[t];
    else if (t < 12) bnbuf[t] = b1[t - 9];
}

int pt_grid(int n)
{
    int g = (n + PT_WARPS - 1) / PT_WARPS;
    const int cap = 148 * 4;
    return g < 1 ? 1 : (g > cap ? cap : g);
}

extern "C" int cb_pt_rel(int n, int k, const float *p, const int *idx, float *rel, double *moments, void *stream)
{
    CB_REQUIRE(n >= 0 && k > 0 && p && idx && rel && moments, CB_EINVAL, "cb_pt_rel: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(moments, 0, 9 * sizeof(double), st);
    const long long rows = (long long)n * k;
    if (rows > 0) {
        int g = (int)((rows + PT_THREADS - 1) / PT_THREADS);
        if (g > 148 * 8) g = 148 * 8;
        k_pt_rel<<<g, PT_THREADS, 0, st>>>(p, idx, rows, k, rel, moments);
    }
    CB_COUNT(2);
    CB_CUDA_CHECK("cb_pt_rel");
    return CB_OK;
}

// F3, staged variant (k <= 16): a block owns PB = 256 / CS points.  v = relu(bn3(w2)) of those points is staged ONCE in
// shared memory (the generic kernel above re-derives it CS times per element), thread (point, j) keeps its row of W4 and
// its k logits in registers and reads v with broadcast LDS.128.
template <int CS>
__global__ void __launch_bounds__(256) k_pt_softmax_staged(int n, int k, const float *__restrict__ w2, const float *__restrict__ bn3,
                                                           const float *__restrict__ w4, const float *__restrict__ b4,
                                                           float *__restrict__ a)
{
    constexpr int PB = 256 / CS;
    extern __shared__ __align__(16) float ss_sm[];             // [PB][k][CS]
    const int tp = threadIdx.x / CS, tj = threadIdx.x % CS;
    float wrow[CS];
#pragma unroll
    for (int i = 0; i < CS; i++) wrow[i] = __ldg(w4 + tj * CS + i);
    const float bj = __ldg(b4 + tj);
    const int tile_elems = PB * k * CS;
    const int tiles = (n + PB - 1) / PB;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long base = (long long)tile * tile_elems;
        const long long limit = (long long)n * k * CS;
        for (int e = threadIdx.x; e < tile_elems; e += 256) {
            const int i = e % CS;
            float v = 0.f;
            if (base + e < limit) v = fmaxf(__ldg(w2 + base + e) * __ldg(bn3 + i) + __ldg(bn3 + CS + i), 0.f);
            ss_sm[e] = v;
        }
        __syncthreads();
        const int pt = tile * PB + tp;
        if (pt < n) {
            float lg[16];
            float mx = -3.0e38f;
#pragma unroll
            for (int kk = 0; kk < 16; kk++) {
                lg[kk] = -3.0e38f;
                if (kk < k) {
                    const float *vr = ss_sm + (tp * k + kk) * CS;
                    float acc = bj;
#pragma unroll
                    for (int i = 0; i < CS; i += 4) {
                        const float4 v4 = *reinterpret_cast<const float4 *>(vr + i);
                        acc += wrow[i] * v4.x; acc += wrow[i + 1] * v4.y; acc += wrow[i + 2] * v4.z; acc += wrow[i + 3] * v4.w;
                    }
                    lg[kk] = acc;
                    mx = fmaxf(mx, acc);
                }
            }
            float sum = 0.f;
#pragma unroll
            for (int kk = 0; kk < 16; kk++)
                if (kk < k) { lg[kk] = expf(lg[kk] - mx); sum += lg[kk]; }
            const float inv = 1.0f / sum;
#pragma unroll
            for (int kk = 0; kk < 16; kk++)
                if (kk < k) a[((size_t)pt * k + kk) * CS + tj] = lg[kk] * inv;
        }
        __syncthreads();
    }
}

template <int CS>
static void pt_softmax_staged_launch(int n, int k, const float *w2, const float *bn3, const float *w4, const float *b4, float *a,
                                     cudaStream_t st)
{
    constexpr int PB = 256 / CS;
    const size_t smem = (size_t)PB * k * CS * sizeof(float);       // 256 * k floats <= 16 KB
    int tiles = (n + PB - 1) / PB;
    int g = tiles < 148 * 6 ? (tiles < 1 ? 1 : tiles) : 148 * 6;
    k_pt_softmax_staged<CS><<<g, 256, smem, st>>>(n, k, w2, bn3, w4, b4, a);
}

int cb_pt_mma_enabled();
void cb_pt_w2_mma(int c, int n, int k, int ld, const float *rel, const int *idx, const float *xq, const float *xk,
                  const float *w2p, const float *b2p, const float *smalld, const float *bn2, const float *w3, const float *b3,
                  float *w2out, double *stats, const float *w0, cudaStream_t st);

template <int C>
static int pt_forward_c(int n, int k, int ld, const CbPtLayer *L, const float *rel, const double *moments, const int *idx,
                        const float *xq, const float *xk, const float *xv, float *out, float *w2buf, float *abuf,
                        float *bnbuf, double *stats, float *w0buf, cudaStream_t st)
{
    constexpr int CS = C / 8;
    float *small = bnbuf, *bn2 = bnbuf + 24, *bn3 = bnbuf + 24 + 4 * C;
    double *stats2 = stats, *stats3 = stats + 2 * C;
    const double rows = (double)n * (double)k;
    cudaMemsetAsync(stats, 0, sizeof(double) * (2 * C + 2 * CS), st);
    k_pt_pack_small<<<1, 32, 0, st>>>(L->w1, L->b1, bnbuf);
    k_pt_bn1_prep<<<1, 32, 0, st>>>(moments, rows, L->w1, L->b1, L->bn1_weight, L->bn1_bias, L->bn1_running_mean,
                                     L->bn1_running_var, L->momentum, L->eps, L->training, bnbuf + 12);
    const int grid = pt_grid(n);
    if (L->training)
        k_pt_w0_stats<C><<<grid, PT_THREADS, 0, st>>>(n, k, ld, rel, idx, xq, xk, L->w2, L->b2, small, stats2, w0buf);
    k_bn_finalize<<<(C + 127) / 128, 128, 0, st>>>(stats2, rows, C, L->bn2_weight, L->bn2_bias, L->bn2_running_mean,
                                                    L->bn2_running_var, L->momentum, L->eps, L->training, bn2);
    if (cb_pt_mma_enabled() && ld % 4 == 0) {
        // tensor cores (3xTF32), ptlayer_mma.cu
        cb_pt_w2_mma(C, n, k, ld, rel, idx, xq, xk, L->w2, L->b2, small, bn2, L->w3, L->b3, w2buf, stats3,
                     L->training ? w0buf : nullptr, st);
    } else {
        const size_t smem = (size_t)CS * C * sizeof(float);
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_pt_w2<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_pt_w2<C><<<grid, PT_THREADS, smem, st>>>(n, k, ld, rel, idx, xq, xk, L->w2, L->b2, small, bn2, L->w3, L->b3, w2buf, stats3);
    }
    k_bn_finalize<<<1, 128, 0, st>>>(stats3, rows, CS, L->bn3_weight, L->bn3_bias, L->bn3_running_mean,
                                      L->bn3_running_var, L->momentum, L->eps, L->training, bn3);
    const size_t smem4 = (size_t)(CS * (CS + 1) + 2 * CS) * sizeof(float);
    long long tot = (long long)n * CS;
    int g4 = (int)((tot + 255) / 256);
    if (g4 > 148 * 8) g4 = 148 * 8;
    if (g4 < 1) g4 = 1;
    if (k <= 16) pt_softmax_staged_launch<CS>(n, k, w2buf, bn3, L->w4, L->b4, abuf, st);
    else k_pt_softmax<<<g4, 256, smem4, st>>>(n, k, CS, w2buf, bn3, L->w4, L->b4, abuf);
    k_pt_aggregate<C><<<grid, PT_THREADS, 0, st>>>(n, k, ld, rel, idx, xv, L->w2, L->b2, small, abuf, out);
    CB_COUNT(9);
    CB_CUDA_CHECK("cb_pt_layer_forward");
    return CB_OK;
}

extern "C" size_t cb_pt_bnbuf_floats(int c) { return (size_t)(24 + 4 * c + 4 * (c / 8)); }
extern "C" size_t cb_pt_stats_doubles(int c) { return (size_t)(2 * c + 2 * (c / 8) + 64); }

extern "C" int cb_pt_layer_forward(int n, int k, int c, int ld, const CbPtLayer *L, const float *rel, const double *moments,
                                   const int *idx, const float *xq, const float *xk, const float *xv, float *out,
                                   float *w2buf, float *abuf, float *bnbuf, double *stats, float *w0buf, void *stream)
{
    CB_REQUIRE(n >= 0 && k >= 1 && k <= PT_KMAX, CB_EINVAL, "cb_pt_layer_forward: n=%d k=%d (k <= %d)", n, k, PT_KMAX);
    CB_REQUIRE(ld >= c && ld % 4 == 0, CB_EINVAL, "cb_pt_layer_forward: ld=%d (row stride of x_q/x_k/x_v) must be >= c and a multiple of 4", ld);
    CB_REQUIRE(L && rel && moments && idx && xq && xk && xv && out && w2buf && abuf && bnbuf && stats, CB_EINVAL,
               "cb_pt_layer_forward: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return CB_OK;
    switch (c) {
    case 32: return pt_forward_c<32>(n, k, ld, L, rel, moments, idx, xq, xk, xv, out, w2buf, abuf, bnbuf, stats, w0buf, st);
    case 64: return pt_forward_c<64>(n, k, ld, L, rel, moments, idx, xq, xk, xv, out, w2buf, abuf, bnbuf, stats, w0buf, st);
    case 128: return pt_forward_c<128>(n, k, ld, L, rel, moments, idx, xq, xk, xv, out, w2buf, abuf, bnbuf, stats, w0buf, st);
    case 256: return pt_forward_c<256>(n, k, ld, L, rel, moments, idx, xq, xk, xv, out, w2buf, abuf, bnbuf, stats, w0buf, st);
    case 512: return pt_forward_c<512>(n, k, ld, L, rel, moments, idx, xq, xk, xv, out, w2buf, abuf, bnbuf, stats, w0buf, st);
    default:
        cb_set_error("cb_pt_layer_forward: c=%d unsupported (32,64,128,256,512)", c);
        return CB_EUNSUPPORTED;
    }
}
