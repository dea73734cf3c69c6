// fps.cu — a2: farthest point sampling, bit-identical to the reference
// (pytorch/lib/pointops/src/sampling/sampling_cuda_kernel.cu:14-171).
//
// The reference runs ONE block per scene and re-reads xyz and the running min-distance buffer
// from global memory in each of its m-1 strictly sequential iterations.  Here each scene is owned
// by a thread-block CLUSTER: every thread keeps its points and their running min-distance in
// registers for the whole kernel; one iteration = per-thread update -> warp arg-max (REDUX) ->
// CTA arg-max through shared memory -> cluster arg-max through distributed shared memory.
//
// Tie rule.  The reference's winner among equal maxima is fixed by its launch geometry: thread
// t = r % B of a B-thread block (B = opt_n_threads(n_max), cuda_utils.h:11-14) scans r = t, t+B, ...
// keeping the first strict maximum, and the shared-memory tree keeps the lower slot unless the upper
// is strictly greater (sampling_cuda_kernel.cu:5-10) — so the winner has the smallest
// (bit-reverse_B(t), r / B).  We give every point that 32-bit priority and reduce
// the pair (distance bits, ~priority) with max, which reproduces the rule under any reduction order.
#include "knn.cuh"
#include <cooperative_groups.h>
#include <math.h>
namespace cg = cooperative_groups;

#define FPS_THREADS 1024
static int g_fps_mode = 0;
static int g_fps_ws_min = 8192;

struct FpsCand {
    unsigned d;   // distance bits (non-negative floats order like unsigned ints)
    unsigned np;  // ~priority (larger wins)
};

__device__ __forceinline__ bool fps_better(unsigned d1, unsigned p1, unsigned d0, unsigned p0)
{
    return d1 > d0 || (d1 == d0 && p1 > p0);
}

template <int CL, int PPT>
__global__ void __launch_bounds__(FPS_THREADS, 1) k_fps(const float *__restrict__ xyz, const int *__restrict__ offset,
                                                        const int *__restrict__ new_offset, float *__restrict__ tmp,
                                                        int *__restrict__ idx, int logB)
{
    __shared__ FpsCand s_warp[2][32];
    __shared__ FpsCand s_xchg[2][CL > 1 ? CL : 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int rank = 0;
    if (CL > 1) rank = (int)cg::this_cluster().block_rank();
    const int bid = blockIdx.x / CL;
    const int start_n = bid == 0 ? 0 : offset[bid - 1], end_n = offset[bid];
    const int start_m = bid == 0 ? 0 : new_offset[bid - 1], end_m = new_offset[bid];
    const int ns = end_n - start_n;
    const int B = 1 << logB;
    const unsigned S = (unsigned)((ns + B - 1) >> logB);     // scan steps per reference thread

    constexpr int NP = PPT > 0 ? PPT : 1;
    float px[NP], py[NP], pz[NP], md[NP];
    unsigned npri[NP];
    const int stride = CL * FPS_THREADS;
    const int r0 = rank * FPS_THREADS + tid;
    if (PPT > 0) {
#pragma unroll
        for (int u = 0; u < NP; u++) {
            const int r = r0 + u * stride;
            if (r < ns) {
                const int k = start_n + r;
                px[u] = xyz[3 * k]; py[u] = xyz[3 * k + 1]; pz[u] = xyz[3 * k + 2];
                md[u] = tmp[k];
                const unsigned t = (unsigned)r & (unsigned)(B - 1);
                const unsigned brev = logB ? (__brev(t) >> (32 - logB)) : 0u;
                npri[u] = ~(brev * S + ((unsigned)r >> logB));
            } else {
                px[u] = py[u] = pz[u] = 0.f; md[u] = -1.f; npri[u] = 0u;
            }
        }
    }
    if (ns <= 0 || end_m <= start_m) return;   // uniform per cluster
    if (rank == 0 && tid == 0) idx[start_m] = start_n;       // sampling_cuda_kernel.cu:39
    int old = start_n;
    for (int j = start_m + 1; j < end_m; j++) {
        const int par = j & 1;
        const float ox = __ldg(xyz + 3 * old), oy = __ldg(xyz + 3 * old + 1), oz = __ldg(xyz + 3 * old + 2);
        unsigned bd = 0u, bp = 0u;    // (0, 0) loses to every real point (np of real points is >= ~(B*S) > 0)
        if (PPT > 0) {
#pragma unroll
            for (int u = 0; u < NP; u++) {
                const float d = cb_sqdist(px[u], py[u], pz[u], ox, oy, oz);
                const float m2 = fminf(d, md[u]);
                md[u] = m2;
                const unsigned db = __float_as_uint(fmaxf(m2, 0.f));
                const bool real = npri[u] != 0u;
                if (real && fps_better(db, npri[u], bd, bp)) { bd = db; bp = npri[u]; }
            }
        } else {
            for (int r = r0; r < ns; r += stride) {
                const int k = start_n + r;
                const float d = cb_sqdist(xyz[3 * k], xyz[3 * k + 1], xyz[3 * k + 2], ox, oy, oz);
                const float m2 = fminf(d, tmp[k]);
                tmp[k] = m2;
                const unsigned t = (unsigned)r & (unsigned)(B - 1);
                const unsigned brev = logB ? (__brev(t) >> (32 - logB)) : 0u;
                const unsigned np = ~(brev * S + ((unsigned)r >> logB));
                const unsigned db = __float_as_uint(fmaxf(m2, 0.f));
                if (fps_better(db, np, bd, bp)) { bd = db; bp = np; }
            }
        }
        // warp arg-max
        unsigned wd = __reduce_max_sync(CB_FULL_MASK, bd);
        unsigned wp = __reduce_max_sync(CB_FULL_MASK, bd == wd ? bp : 0u);
        if (lane == 0) { s_warp[par][warp].d = wd; s_warp[par][warp].np = wp; }
        __syncthreads();
        // CTA arg-max (every warp redundantly)
        const FpsCand c = s_warp[par][lane];
        unsigned cd = __reduce_max_sync(CB_FULL_MASK, c.d);
        unsigned cp = __reduce_max_sync(CB_FULL_MASK, c.d == cd ? c.np : 0u);
        if (CL > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            if (warp == 0 && lane < CL) {
                FpsCand *dst = cluster.map_shared_rank(&s_xchg[par][rank], lane);
                dst->d = cd; dst->np = cp;
            }
            cluster.sync();
            const FpsCand x = s_xchg[par][lane < CL ? lane : 0];
            cd = __reduce_max_sync(CB_FULL_MASK, x.d);
            cp = __reduce_max_sync(CB_FULL_MASK, x.d == cd ? x.np : 0u);
        }
        // priority -> point index
        const unsigned pri = ~cp;
        const unsigned brev = pri / S, step = pri - brev * S;
        const unsigned t = logB ? (__brev(brev) >> (32 - logB)) : 0u;
        old = start_n + (int)((step << logB) + t);
        if (rank == 0 && tid == 0) idx[j] = old;
    }
    if (PPT > 0) {   // leave the running min-distance where the reference leaves it
#pragma unroll
        for (int u = 0; u < NP; u++) {
            const int r = r0 + u * stride;
            if (r < ns) tmp[start_n + r] = md[u];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Bucket-pruned exact FPS (scenes whose running min-distance array fits in shared memory).
//
// The supports are first sorted by grid cell (the KNN grid of knn.cu), so 32 consecutive sorted
// points form a spatially compact BUCKET with a small bounding box.  An iteration only has to touch
// buckets whose box is closer to the new sample than the bucket's current maximum min-distance —
// for every other bucket min(d, tmp) == tmp for all its points, so skipping them is exact.  After the
// first few dozen samples only the handful of buckets around the new sample are touched, which turns
// the reference's O(n) sweep per iteration into O(n / j).
// One CTA of 32 warps per scene.  Bucket b is OWNED by lane (b/32)%32 of warp b%32 (slot b/1024): its
// box, its current arg-max (distance bits, ~priority, coordinates, index) live in that lane's
// registers; the min-distances live in shared memory.  Per iteration: register-only box tests ->
// the warp refreshes its touched buckets (one 32-point load each) -> warp arg-max -> ONE
// __syncthreads -> every warp reduces the 32 warp results.  Tie rule: see the header of this file.
// ---------------------------------------------------------------------------------------------
#define FPSB_SLOTS 2            // buckets per lane  -> up to 2048 buckets = 65536 points per scene
#define FPSB_MAX_POINTS 49152   // 192 KB of min-distances in shared memory

__device__ unsigned long long g_fps_dbg[8];   // [0] touched buckets, [1] iterations, [2] cycles (warp 0), [3] cycles in refresh (warp 0)
__device__ int g_fps_dbg_on = 0;

struct FpsSlot {
    unsigned d, np;
    float x, y, z;
    int orig;
};

__global__ void __launch_bounds__(1024, 1) k_fps_bucket(const float4 *__restrict__ sorted, const float *__restrict__ xyz,
                                                        const int *__restrict__ offset, const int *__restrict__ new_offset,
                                                        float *__restrict__ tmp, int *__restrict__ idx, int logB)
{
    extern __shared__ float md[];                          // running min distance, sorted order
    __shared__ FpsSlot slots[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bid = blockIdx.x;
    const int S0 = bid == 0 ? 0 : offset[bid - 1], S1 = offset[bid];
    const int start_m = bid == 0 ? 0 : new_offset[bid - 1], end_m = new_offset[bid];
    const int ns = S1 - S0;
    if (ns <= 0 || end_m <= start_m) return;
    const int Bref = 1 << logB;
    const unsigned S = (unsigned)((ns + Bref - 1) >> logB);
    const int nb = (ns + 31) >> 5;

    float lox[FPSB_SLOTS], loy[FPSB_SLOTS], loz[FPSB_SLOTS], hix[FPSB_SLOTS], hiy[FPSB_SLOTS], hiz[FPSB_SLOTS];
    unsigned bd[FPSB_SLOTS], bp[FPSB_SLOTS];
    float bx[FPSB_SLOTS], by[FPSB_SLOTS], bz[FPSB_SLOTS];
    int bo[FPSB_SLOTS];
#pragma unroll
    for (int t = 0; t < FPSB_SLOTS; t++) {
        lox[t] = loy[t] = loz[t] = 3.0e38f; hix[t] = hiy[t] = hiz[t] = -3.0e38f;
        bd[t] = 0u; bp[t] = 0u; bx[t] = by[t] = bz[t] = 0.f; bo[t] = S0;
    }
    auto priority = [&](int orig) -> unsigned {
        const unsigned r = (unsigned)(orig - S0);
        const unsigned tt = r & (unsigned)(Bref - 1);
        const unsigned brev = logB ? (__brev(tt) >> (32 - logB)) : 0u;
        return ~(brev * S + (r >> logB));
    };
    // refresh bucket b against sample (sx,sy,sz); with init=true also builds the box and loads tmp
    auto refresh = [&](int b, int owner, int t, float sx, float sy, float sz, bool init) {
        const int i = (b << 5) + lane;
        const bool valid = i < ns;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) p = __ldg(sorted + S0 + i);
        const int orig = __float_as_int(p.w);
        float m;
        if (init) {
            m = valid ? tmp[orig] : 0.f;
        } else {
            const float d = cb_sqdist(p.x, p.y, p.z, sx, sy, sz);
            m = valid ? fminf(d, md[i]) : 0.f;
        }
        if (valid) md[i] = m;
        const unsigned db = valid ? __float_as_uint(fmaxf(m, 0.f)) : 0u;
        const unsigned np = valid ? priority(orig) : 0u;
        const unsigned wd = __reduce_max_sync(CB_FULL_MASK, db);
        const unsigned wp = __reduce_max_sync(CB_FULL_MASK, db == wd ? np : 0u);
        const int src = __ffs(__ballot_sync(CB_FULL_MASK, db == wd && np == wp)) - 1;
        const float cx = __shfl_sync(CB_FULL_MASK, p.x, src), cy = __shfl_sync(CB_FULL_MASK, p.y, src),
                    cz = __shfl_sync(CB_FULL_MASK, p.z, src);
        const int co = __shfl_sync(CB_FULL_MASK, orig, src);
        float mnx = 0, mny = 0, mnz = 0, mxx = 0, mxy = 0, mxz = 0;
        if (init) {
            const unsigned big = 0xffffffffu;
            mnx = cb_ord2f(__reduce_min_sync(CB_FULL_MASK, valid ? cb_f2ord(p.x) : big));
            mny = cb_ord2f(__reduce_min_sync(CB_FULL_MASK, valid ? cb_f2ord(p.y) : big));
            mnz = cb_ord2f(__reduce_min_sync(CB_FULL_MASK, valid ? cb_f2ord(p.z) : big));
            mxx = cb_ord2f(__reduce_max_sync(CB_FULL_MASK, valid ? cb_f2ord(p.x) : 0u));
            mxy = cb_ord2f(__reduce_max_sync(CB_FULL_MASK, valid ? cb_f2ord(p.y) : 0u));
            mxz = cb_ord2f(__reduce_max_sync(CB_FULL_MASK, valid ? cb_f2ord(p.z) : 0u));
        }
        if (lane == owner) {
#pragma unroll
            for (int tt = 0; tt < FPSB_SLOTS; tt++)
                if (tt == t) {
                    bd[tt] = wd; bp[tt] = wp; bx[tt] = cx; by[tt] = cy; bz[tt] = cz; bo[tt] = co;
                    if (init) { lox[tt] = mnx; loy[tt] = mny; loz[tt] = mnz; hix[tt] = mxx; hiy[tt] = mxy; hiz[tt] = mxz; }
                }
        }
    };
    // ---- init: every warp builds its own buckets: b = warp + 32 * (owner + 32 * t)
#pragma unroll 1
    for (int t = 0; t < FPSB_SLOTS; t++)
#pragma unroll 1
        for (int owner = 0; owner < 32; owner++) {
            const int b = warp + 32 * (owner + 32 * t);
            if (b < nb) refresh(b, owner, t, 0.f, 0.f, 0.f, true);
        }
    int old = S0;
    float sx = __ldg(xyz + 3 * old), sy = __ldg(xyz + 3 * old + 1), sz = __ldg(xyz + 3 * old + 2);
    if (tid == 0) idx[start_m] = old;                        // sampling_cuda_kernel.cu:39
    const int dbg = g_fps_dbg_on;
    long long t_start = 0, t_refresh = 0;
    unsigned long long n_touched = 0;
    if (dbg) t_start = clock64();
    for (int j = start_m + 1; j < end_m; j++) {
        const int par = j & 1;
        // 1. which of my buckets can change?  (box distance, with a 1e-4 safety margin for fp32 rounding)
#pragma unroll
        for (int t = 0; t < FPSB_SLOTS; t++) {
            const float dx = fmaxf(fmaxf(lox[t] - sx, sx - hix[t]), 0.f), dy = fmaxf(fmaxf(loy[t] - sy, sy - hiy[t]), 0.f),
                        dz = fmaxf(fmaxf(loz[t] - sz, sz - hiz[t]), 0.f);
            const float lb = (dx * dx + dy * dy + dz * dz) * 0.9999f;
            unsigned touched = __ballot_sync(CB_FULL_MASK, bd[t] != 0u && lb < __uint_as_float(bd[t]));
            if (dbg) n_touched += __popc(touched);
            const long long c0 = dbg ? clock64() : 0;
            while (touched) {
                const int owner = __ffs(touched) - 1;
                touched &= touched - 1;
                refresh(warp + 32 * (owner + 32 * t), owner, t, sx, sy, sz, false);
            }
            if (dbg) t_refresh += clock64() - c0;
        }
        // 2. warp arg-max over the buckets its lanes own
        unsigned md_ = bd[0], mp_ = bp[0];
        int mt = 0;
#pragma unroll
        for (int t = 1; t < FPSB_SLOTS; t++)
            if (bd[t] > md_ || (bd[t] == md_ && bp[t] > mp_)) { md_ = bd[t]; mp_ = bp[t]; mt = t; }
        float mx = bx[0], my = by[0], mz = bz[0];
        int mo = bo[0];
#pragma unroll
        for (int t = 1; t < FPSB_SLOTS; t++)
            if (mt == t) { mx = bx[t]; my = by[t]; mz = bz[t]; mo = bo[t]; }
        {
            const unsigned wd = __reduce_max_sync(CB_FULL_MASK, md_);
            const unsigned wp = __reduce_max_sync(CB_FULL_MASK, md_ == wd ? mp_ : 0u);
            const int src = __ffs(__ballot_sync(CB_FULL_MASK, md_ == wd && mp_ == wp)) - 1;
            if (lane == src) {
                FpsSlot sl;
                sl.d = wd; sl.np = wp; sl.x = mx; sl.y = my; sl.z = mz; sl.orig = mo;
                slots[par][warp] = sl;
            }
        }
        __syncthreads();
        // 3. CTA arg-max (every warp redundantly)
        {
            const FpsSlot sl = slots[par][lane];
            const unsigned wd = __reduce_max_sync(CB_FULL_MASK, sl.d);
            const unsigned wp = __reduce_max_sync(CB_FULL_MASK, sl.d == wd ? sl.np : 0u);
            const int src = __ffs(__ballot_sync(CB_FULL_MASK, sl.d == wd && sl.np == wp)) - 1;
            sx = __shfl_sync(CB_FULL_MASK, sl.x, src); sy = __shfl_sync(CB_FULL_MASK, sl.y, src);
            sz = __shfl_sync(CB_FULL_MASK, sl.z, src);
            old = __shfl_sync(CB_FULL_MASK, sl.orig, src);
        }
        if (tid == 0) idx[j] = old;
    }
    if (dbg && lane == 0) {
        atomicAdd(&g_fps_dbg[0], n_touched);
        if (warp == 0 && bid == 0) {
            g_fps_dbg[1] = (unsigned long long)(end_m - start_m - 1);
            g_fps_dbg[2] = (unsigned long long)(clock64() - t_start);
            g_fps_dbg[3] = (unsigned long long)t_refresh;
        }
    }
    // leave the running min-distance where the reference leaves it
    __syncthreads();
    for (int i = tid; i < ns; i += 1024) tmp[__float_as_int(__ldg(sorted + S0 + i).w)] = md[i];
}

// ---------------------------------------------------------------------------------------------
// Cluster variant of the bucket-pruned FPS: one thread-block CLUSTER of CL CTAs per scene.
//
// The single-CTA kernel above runs 32 warps that all execute the same serial chain (box tests -> refresh ->
// warp arg-max -> __syncthreads -> CTA arg-max) — eight warps per scheduler compete for issue slots, every
// refreshed bucket is a 32-point load from global memory (L2 latency), and the arg-max has two levels.
// Here each CTA owns a contiguous 1/CL of the cell-sorted points and keeps BOTH their coordinates and their
// running min-distances in its shared memory (20 B per point), runs only W = 4 (or 8) warps — one or two per
// scheduler, a lane owns up to S buckets — and the arg-max is ONE all-to-all step: every warp of the cluster sends
// its candidate (distance bits, ~priority, x, y, z, index: 24 B) straight into the slot arrays of all CL CTAs with
// st.async (distributed shared memory), whose arrival is counted by the receiver's mbarrier (complete_tx); each
// warp then reduces the CL*W slots redundantly.  No __syncthreads, no cluster barrier and no global-memory access
// inside the iteration.  Tie rule: identical to the kernels above (priority = reference thread/stride order).
// ---------------------------------------------------------------------------------------------
#define FPSC_MAX_PER_CTA 10752

__device__ __forceinline__ unsigned fps_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned fps_mapa(unsigned addr, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}

template <int CL, int W, int S>
__global__ void __launch_bounds__(W * 32, 1) k_fps_bucket_cl(const float4 *__restrict__ sorted, const float *__restrict__ xyz,
                                                             const int *__restrict__ offset, const int *__restrict__ new_offset,
                                                             float *__restrict__ tmp, int *__restrict__ idx, int logB, int cap)
{
    constexpr int NT = W * 32, NSLOT = CL * W;
    static_assert(NSLOT <= 64, "slot reduction handles at most 64 warps per cluster");
    extern __shared__ __align__(16) unsigned char fps_sm[];
    float4 *spt = reinterpret_cast<float4 *>(fps_sm);            // cap points (x, y, z, original index bits)
    float *md = reinterpret_cast<float *>(spt + cap);            // cap running min distances
    __shared__ __align__(16) uint4 slotA[2][NSLOT];              // (d bits, ~priority, x, y)
    __shared__ __align__(8) uint2 slotB[2][NSLOT];               // (z, original index)
    __shared__ __align__(8) unsigned long long bar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int bid = blockIdx.x / CL;
    const int S0 = bid == 0 ? 0 : offset[bid - 1], S1 = offset[bid];
    const int start_m = bid == 0 ? 0 : new_offset[bid - 1], end_m = new_offset[bid];
    const int ns = S1 - S0;
    const int Bref = 1 << logB;
    const unsigned SS = (unsigned)((ns + Bref - 1) >> logB);
    const int base = (int)rank * cap;                            // my chunk of the sorted order: [base, base + cnt)
    const int cnt = max(0, min(cap, ns - base));
    const int nbl = (cnt + 31) >> 5;                             // my buckets (<= W * 32 * S)

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fps_smem_u32(&bar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fps_smem_u32(&bar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < cnt; i += NT) {                        // stage my points
        const float4 p = __ldg(sorted + S0 + base + i);
        spt[i] = p;
        md[i] = tmp[__float_as_int(p.w)];
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");

    float lox[S], loy[S], loz[S], hix[S], hiy[S], hiz[S];
    unsigned bd[S], bp[S];
    float bx[S], by[S], bz[S];
    int bo[S];
#pragma unroll
    for (int t = 0; t < S; t++) {
        lox[t] = loy[t] = loz[t] = 3.0e38f; hix[t] = hiy[t] = hiz[t] = -3.0e38f;
        bd[t] = 0u; bp[t] = 0u; bx[t] = by[t] = bz[t] = 0.f; bo[t] = S0;
    }
    auto priority = [&](int orig) -> unsigned {
        const unsigned r = (unsigned)(orig - S0);
        const unsigned tt = r & (unsigned)(Bref - 1);
        const unsigned brev = logB ? (__brev(tt) >> (32 - logB)) : 0u;
        return ~(brev * SS + (r >> logB));
    };
    // refresh local bucket bl (owned by slot t of lane `owner`) against sample (sx,sy,sz); init: build the box instead
    auto refresh = [&](int bl, int owner, int t, float sx, float sy, float sz, bool init) {
        const int i = (bl << 5) + lane;
        const bool valid = i < cnt;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) p = spt[i];
        const int orig = __float_as_int(p.w);
        float m = 0.f;
        if (valid) {
            m = md[i];
            if (!init) {
                m = fminf(cb_sqdist(p.x, p.y, p.z, sx, sy, sz), m);
                md[i] = m;
            }
        }
        const unsigned db = valid ? __float_as_uint(fmaxf(m, 0.f)) : 0u;
        const unsigned np = valid ? priority(orig) : 0u;
        const unsigned wd = __reduce_max_sync(CB_FULL_MASK, db);
        const unsigned wp = __reduce_max_sync(CB_FULL_MASK, db == wd ? np : 0u);
        const int src = __ffs(__ballot_sync(CB_FULL_MASK, db == wd && np == wp)) - 1;
        const float cx = __shfl_sync(CB_FULL_MASK, p.x, src), cy = __shfl_sync(CB_FULL_MASK, p.y, src),
                    cz = __shfl_sync(CB_FULL_MASK, p.z, src);
        const int co = __shfl_sync(CB_FULL_MASK, orig, src);
        float mnx = 0, mny = 0, mnz = 0, mxx = 0, mxy = 0, mxz = 0;
        if (init) {
            const unsigned big = 0xffffffffu;
            mnx = cb_ord2f(__reduce_min_sync(CB_FULL_MASK, valid ? cb_f2ord(p.x) : big));
            mny = cb_ord2f(__reduce_min_sync(CB_FULL_MASK, valid ? cb_f2ord(p.y) : big));
            mnz = cb_ord2f(__reduce_min_sync(CB_FULL_MASK, valid ? cb_f2ord(p.z) : big));
            mxx = cb_ord2f(__reduce_max_sync(CB_FULL_MASK, valid ? cb_f2ord(p.x) : 0u));
            mxy = cb_ord2f(__reduce_max_sync(CB_FULL_MASK, valid ? cb_f2ord(p.y) : 0u));
            mxz = cb_ord2f(__reduce_max_sync(CB_FULL_MASK, valid ? cb_f2ord(p.z) : 0u));
        }
        if (lane == owner) {
#pragma unroll
            for (int tt = 0; tt < S; tt++)
                if (tt == t) {
                    bd[tt] = wd; bp[tt] = wp; bx[tt] = cx; by[tt] = cy; bz[tt] = cz; bo[tt] = co;
                    if (init) { lox[tt] = mnx; loy[tt] = mny; loz[tt] = mnz; hix[tt] = mxx; hiy[tt] = mxy; hiz[tt] = mxz; }
                }
        }
    };
    // local bucket bl = warp + W * (owner + 32 * t) is owned by slot t of lane `owner` of warp `warp`
#pragma unroll 1
    for (int t = 0; t < S; t++)
#pragma unroll 1
        for (int owner = 0; owner < 32; owner++) {
            const int bl = warp + W * (owner + 32 * t);
            if (bl < nbl) refresh(bl, owner, t, 0.f, 0.f, 0.f, true);
        }
    int old = S0;
    float sx = __ldg(xyz + 3 * old), sy = __ldg(xyz + 3 * old + 1), sz = __ldg(xyz + 3 * old + 2);
    if (rank == 0 && tid == 0 && end_m > start_m) idx[start_m] = old;      // sampling_cuda_kernel.cu:39
    // remote addresses of my warp's slot and of the barriers in CTA `lane` (lanes < CL send)
    const unsigned my_slot = rank * (unsigned)W + (unsigned)warp;
    unsigned rA = 0, rB = 0, rbar = 0;
    if (lane < CL) {
        rA = fps_mapa(fps_smem_u32(&slotA[0][my_slot]), (unsigned)lane);
        rB = fps_mapa(fps_smem_u32(&slotB[0][my_slot]), (unsigned)lane);
        rbar = fps_mapa(fps_smem_u32(&bar[0]), (unsigned)lane);
    }
    const unsigned bar_local = fps_smem_u32(&bar[0]);
    const int iters = (ns > 0 && end_m > start_m) ? end_m - start_m - 1 : 0;     // uniform across the cluster
    const int dbg = g_fps_dbg_on;
    long long t_start = 0, t_refresh = 0, t_xchg = 0;
    unsigned long long n_touched = 0;
    unsigned max_touched = 0;
    if (dbg) t_start = clock64();
    for (int it = 0; it < iters; it++) {
        const int par = it & 1;
        // arm this iteration's barrier first (its phase cannot complete before this arrival, however early the
        // candidates of faster warps land)
        if (tid == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_local + (unsigned)par * 8u),
                         "r"((unsigned)(NSLOT * 24)) : "memory");
        // 1. which of my buckets can change?  (box distance, with a 1e-4 safety margin for fp32 rounding)
        const long long c0 = dbg ? clock64() : 0;
#pragma unroll
        for (int t = 0; t < S; t++) {
            const float dx = fmaxf(fmaxf(lox[t] - sx, sx - hix[t]), 0.f), dy = fmaxf(fmaxf(loy[t] - sy, sy - hiy[t]), 0.f),
                        dz = fmaxf(fmaxf(loz[t] - sz, sz - hiz[t]), 0.f);
            const float lb = (dx * dx + dy * dy + dz * dz) * 0.9999f;
            unsigned touched = __ballot_sync(CB_FULL_MASK, bd[t] != 0u && lb < __uint_as_float(bd[t]));
            if (dbg) { n_touched += __popc(touched); max_touched = max(max_touched, (unsigned)__popc(touched)); }
            while (touched) {
                const int owner = __ffs(touched) - 1;
                touched &= touched - 1;
                refresh(warp + W * (owner + 32 * t), owner, t, sx, sy, sz, false);
            }
        }
        if (dbg) t_refresh += clock64() - c0;
        const long long c1 = dbg ? clock64() : 0;
        // 2. warp arg-max over the buckets its lanes own
        {
            unsigned md_ = bd[0], mp_ = bp[0];
            float mx = bx[0], my = by[0], mz = bz[0];
            int mo = bo[0];
#pragma unroll
            for (int t = 1; t < S; t++)
                if (bd[t] > md_ || (bd[t] == md_ && bp[t] > mp_)) { md_ = bd[t]; mp_ = bp[t]; mx = bx[t]; my = by[t]; mz = bz[t]; mo = bo[t]; }
            const unsigned wd = __reduce_max_sync(CB_FULL_MASK, md_);
            const unsigned wp = __reduce_max_sync(CB_FULL_MASK, md_ == wd ? mp_ : 0u);
            const int src = __ffs(__ballot_sync(CB_FULL_MASK, md_ == wd && mp_ == wp)) - 1;
            const unsigned cx = __float_as_uint(__shfl_sync(CB_FULL_MASK, mx, src)), cy = __float_as_uint(__shfl_sync(CB_FULL_MASK, my, src)),
                           cz = __float_as_uint(__shfl_sync(CB_FULL_MASK, mz, src));
            const unsigned co = (unsigned)__shfl_sync(CB_FULL_MASK, mo, src);
            // 3. all-to-all: lane r sends this warp's candidate to CTA r
            if (lane < CL) {
                const unsigned a = rA + (unsigned)par * (unsigned)sizeof(slotA[0]);
                const unsigned b2 = rB + (unsigned)par * (unsigned)sizeof(slotB[0]);
                const unsigned mb = rbar + (unsigned)par * 8u;
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                             ::"r"(a), "r"(wd), "r"(wp), "r"(cx), "r"(cy), "r"(mb) : "memory");
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
                             ::"r"(b2), "r"(cz), "r"(co), "r"(mb) : "memory");
            }
        }
        // 4. wait for the CL*W candidates of this iteration
        {
            const unsigned mb = bar_local + (unsigned)par * 8u, parity = (unsigned)(it >> 1) & 1u;
            unsigned ok;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(mb), "r"(parity) : "memory");
            } while (!ok);
        }
        if (dbg) t_xchg += clock64() - c1;
        // 5. cluster arg-max (every warp of every CTA redundantly, on its own copy of the slots)
        {
            uint4 best = make_uint4(0u, 0u, 0u, 0u);
            uint2 bestB = make_uint2(0u, (unsigned)S0);
            if (lane < NSLOT) { best = slotA[par][lane]; bestB = slotB[par][lane]; }
            if (NSLOT > 32 && lane + 32 < NSLOT) {
                const uint4 t = slotA[par][lane + 32];
                const uint2 tb = slotB[par][lane + 32];
                if (t.x > best.x || (t.x == best.x && t.y > best.y)) { best = t; bestB = tb; }
            }
            const unsigned wd = __reduce_max_sync(CB_FULL_MASK, best.x);
            const unsigned wp = __reduce_max_sync(CB_FULL_MASK, best.x == wd ? best.y : 0u);
            const int src = __ffs(__ballot_sync(CB_FULL_MASK, best.x == wd && best.y == wp)) - 1;
            sx = __uint_as_float(__shfl_sync(CB_FULL_MASK, best.z, src));
            sy = __uint_as_float(__shfl_sync(CB_FULL_MASK, best.w, src));
            sz = __uint_as_float(__shfl_sync(CB_FULL_MASK, bestB.x, src));
            old = (int)__shfl_sync(CB_FULL_MASK, bestB.y, src);
        }
        if (rank == 0 && tid == 0) idx[start_m + 1 + it] = old;
    }
    if (dbg && lane == 0) {
        atomicAdd(&g_fps_dbg[0], n_touched);
        atomicMax(&g_fps_dbg[5], (unsigned long long)max_touched);
        if (warp == 0 && bid == 0 && rank == 0) {
            g_fps_dbg[1] = (unsigned long long)iters;
            g_fps_dbg[2] = (unsigned long long)(clock64() - t_start);
            g_fps_dbg[3] = (unsigned long long)t_refresh;
            g_fps_dbg[4] = (unsigned long long)t_xchg;
        }
    }
    // leave the running min-distance where the reference leaves it
    __syncthreads();
    for (int i = tid; i < cnt; i += NT) tmp[__float_as_int(spt[i].w)] = md[i];
    // no CTA may exit while a peer could still write into its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CL, int W, int S>
static cudaError_t launch_fps_bucket_cl(int b, int cap, const float4 *sorted, const float *xyz, const int *offset,
                                        const int *new_offset, float *tmp, int *idx, int logB, cudaStream_t st)
{
    const size_t smem = (size_t)cap * 20;
    cudaError_t e = cudaFuncSetAttribute(k_fps_bucket_cl<CL, W, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(b * CL));
    cfg.blockDim = dim3(W * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_fps_bucket_cl<CL, W, S>, sorted, xyz, offset, new_offset, tmp, idx, logB, cap);
}

// slots per lane for `cap` points per CTA and W warps
template <int CL, int W>
static cudaError_t launch_fps_bucket_cl_s(int b, int cap, const float4 *sorted, const float *xyz, const int *offset,
                                          const int *new_offset, float *tmp, int *idx, int logB, cudaStream_t st)
{
    const int buckets = (cap + 31) / 32, s = (buckets + W * 32 - 1) / (W * 32);
    if (s <= 1) return launch_fps_bucket_cl<CL, W, 1>(b, cap, sorted, xyz, offset, new_offset, tmp, idx, logB, st);
    if (s == 2) return launch_fps_bucket_cl<CL, W, 2>(b, cap, sorted, xyz, offset, new_offset, tmp, idx, logB, st);
    if (s == 3) return launch_fps_bucket_cl<CL, W, 3>(b, cap, sorted, xyz, offset, new_offset, tmp, idx, logB, st);
    return cudaErrorInvalidValue;
}

template <int CL, int PPT>
static cudaError_t launch_fps(int b, const float *xyz, const int *offset, const int *new_offset, float *tmp, int *idx,
                              int logB, cudaStream_t st)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(b * CL));
    cfg.blockDim = dim3(FPS_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_fps<CL, PPT>, xyz, offset, new_offset, tmp, idx, logB);
}

extern "C" int cb_furthest_sampling(int b, int n_max, const float *xyz, const int *offset, const int *new_offset,
                                    float *tmp, int *idx, void *stream)
{
    CB_REQUIRE(b >= 0 && n_max >= 0, CB_EINVAL, "cb_furthest_sampling: negative size");
    if (b == 0 || n_max == 0) return CB_OK;
    CB_REQUIRE(xyz && offset && new_offset && tmp && idx, CB_EINVAL, "cb_furthest_sampling: NULL pointer");
    // reference block size: largest power of two <= n_max, capped at 1024 (cuda_utils.h:11-14)
    // (the same double-precision expression, so that exact powers of two round the same way)
    int pow_2 = (int)(log((double)n_max) / log(2.0));
    if (pow_2 > 10) pow_2 = 10;
    if (pow_2 < 0) pow_2 = 0;
    const int logB = pow_2;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (n_max <= 1024) e = launch_fps<1, 1>(b, xyz, offset, new_offset, tmp, idx, logB, st);
    else if (n_max <= 2048) e = launch_fps<1, 2>(b, xyz, offset, new_offset, tmp, idx, logB, st);
    else if (n_max <= 4096) e = launch_fps<1, 4>(b, xyz, offset, new_offset, tmp, idx, logB, st);
    else if (n_max <= 8192) e = launch_fps<1, 8>(b, xyz, offset, new_offset, tmp, idx, logB, st);
    else if (n_max <= 16384) e = launch_fps<8, 2>(b, xyz, offset, new_offset, tmp, idx, logB, st);
    else if (n_max <= 40960) e = launch_fps<8, 5>(b, xyz, offset, new_offset, tmp, idx, logB, st);
    else if (n_max <= 65536) e = launch_fps<8, 8>(b, xyz, offset, new_offset, tmp, idx, logB, st);
    else e = launch_fps<8, 0>(b, xyz, offset, new_offset, tmp, idx, logB, st);
    if (e != cudaSuccess) {
        cb_set_error("cb_furthest_sampling: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return CB_ECUDA;
    }
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_furthest_sampling");
    return CB_OK;
}

// Same operator with a caller-provided workspace (>= cb_knn_workspace_bytes(n, 0, b)): enables the
// bucket-pruned kernel for scenes of 8192 < n_max <= 49152 points.
extern "C" int cb_furthest_sampling_ws(int b, int n_max, const float *xyz, int n, const int *offset,
                                       const int *new_offset, float *tmp, int *idx, void *workspace,
                                       size_t workspace_bytes, void *stream)
{
    if (!workspace || n_max <= g_fps_ws_min || n_max > (g_fps_mode == 1 ? FPSB_MAX_POINTS : 8 * FPSC_MAX_PER_CTA))
        return cb_furthest_sampling(b, n_max, xyz, offset, new_offset, tmp, idx, stream);
    CB_REQUIRE(b > 0 && n >= 0 && xyz && offset && new_offset && tmp && idx, CB_EINVAL, "cb_furthest_sampling_ws: bad arguments");
    CbGridView v;
    const size_t need = cb_grid_layout(n, 0, b, workspace, &v);
    CB_REQUIRE(workspace_bytes >= need, CB_EWORKSPACE, "cb_furthest_sampling_ws: workspace %zu < %zu", workspace_bytes, need);
    CB_REQUIRE(((uintptr_t)workspace & 255) == 0, CB_EINVAL, "cb_furthest_sampling_ws: workspace not 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = cb_grid_build_impl(xyz, n, offset, b, 16, v, st);
    if (rc) return rc;
    int pow_2 = (int)(log((double)n_max) / log(2.0));
    if (pow_2 > 10) pow_2 = 10;
    if (pow_2 < 0) pow_2 = 0;
    if (g_fps_mode != 1) {
        // cluster variant.  mode 0: 8 CTAs x 8 warps | 2: 4 CTAs x 4 warps | 3: 8 CTAs x 4 warps | 4: 4 CTAs x 8 warps
        int cl = (g_fps_mode == 2 || g_fps_mode == 4) ? 4 : 8;
        const int w = (g_fps_mode == 3 || g_fps_mode == 2) ? 4 : 8;
        if ((n_max + cl - 1) / cl > FPSC_MAX_PER_CTA) cl = 8;
        const int cap = (((n_max + cl - 1) / cl) + 31) / 32 * 32;
        if (cap <= FPSC_MAX_PER_CTA) {
            cudaError_t e;
            if (cl == 8 && w == 4) e = launch_fps_bucket_cl_s<8, 4>(b, cap, v.sorted, xyz, offset, new_offset, tmp, idx, pow_2, st);
            else if (cl == 8) e = launch_fps_bucket_cl_s<8, 8>(b, cap, v.sorted, xyz, offset, new_offset, tmp, idx, pow_2, st);
            else if (w == 4) e = launch_fps_bucket_cl_s<4, 4>(b, cap, v.sorted, xyz, offset, new_offset, tmp, idx, pow_2, st);
            else e = launch_fps_bucket_cl_s<4, 8>(b, cap, v.sorted, xyz, offset, new_offset, tmp, idx, pow_2, st);
            if (e != cudaSuccess) {
                cb_set_error("cb_furthest_sampling_ws: %s", cudaGetErrorString(e));
                (void)cudaGetLastError();
                return CB_ECUDA;
            }
            CB_COUNT(1);
            CB_CUDA_CHECK("cb_furthest_sampling_ws");
            return CB_OK;
        }
    }
    CB_REQUIRE(n_max <= FPSB_MAX_POINTS, CB_EUNSUPPORTED, "cb_furthest_sampling_ws: n_max=%d", n_max);
    const size_t smem = (size_t)n_max * sizeof(float);
    cudaFuncSetAttribute(k_fps_bucket, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_fps_bucket<<<b, 1024, smem, st>>>(v.sorted, xyz, offset, new_offset, tmp, idx, pow_2);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_furthest_sampling_ws");
    return CB_OK;
}

// developer knobs: mode 0 = cluster bucket kernel, 8 CTAs x 8 warps (default), 1 = single-CTA bucket kernel,
// 2 / 3 / 4 = cluster kernel with 4x4 / 8x4 / 4x8 (CTAs x warps); ws_min = scenes up to this size use the
// register-resident kernels
extern "C" int cb_fps_set_mode(int mode, int ws_min)
{
    if (mode >= 0 && mode <= 4) g_fps_mode = mode;
    if (ws_min >= 1024) g_fps_ws_min = ws_min;
    return g_fps_mode;
}

extern "C" int cb_debug_fps(int enable, unsigned long long *out8)
{
    if (out8) cudaMemcpyFromSymbol(out8, g_fps_dbg, sizeof(unsigned long long) * 8);
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_fps_dbg, z, sizeof(z));
    cudaMemcpyToSymbol(g_fps_dbg_on, &enable, sizeof(int));
    return 0;
}
