// knn.cuh — uniform-grid K-nearest-neighbour search, device side (shared by knn.cu and the
// fused KNN+gather kernel).  Reference semantics: knnquery_cuda_kernel.cu:65-111.
#pragma once
#include "common.cuh"

struct CbScene {      // one per scene of the support set
    float ox, oy, oz; // grid origin = scene bbox min
    float inv_h, h;   // cell size
    int nx, ny, nz;   // grid dims (each <= CB_GRID_MAX_DIM)
    int cell_base;    // first cell id of this scene in the global cell arrays
    int start, end;   // support point range [start, end)
    int pad;
};

struct CbGridHeader {
    int total_cells;
    int flagged_count;
    int n, b;
    int trial;
    int pad[3];
};

#define CB_GRID_MAX_DIM 1024
#define CB_KNN_MAX_RING 12      // queries needing more rings go to the exact brute-force replay
#define CB_SCAN_TILE 2048

struct CbGridView {   // device pointers into the workspace
    CbGridHeader *hdr;
    CbScene *scenes;
    unsigned *bbox;       // b*6 ordered-uint
    int *occ;             // b*2
    int *tile_sums;
    int *cells;           // counts -> exclusive starts (cell_cap + 1)
    int *coarse;          // trial coarse flags
    int *point_cell;      // n
    int *point_rank;      // n
    float4 *sorted;       // n  (x, y, z, original index bits)
    int *flagged;         // m
    int cell_cap, trial_cap, max_tiles;
};

// fractional cell coordinate; the SAME expression bins supports (count kernel) and queries
__device__ __forceinline__ float cb_cellf(float p, float o, float inv_h) { return __fmul_rn(__fsub_rn(p, o), inv_h); }

size_t cb_grid_layout(int n, int m, int b, void *base, CbGridView *v);
int cb_grid_build_impl(const float *xyz, int n, const int *offset, int b, int nsample_hint, const CbGridView &v,
                       cudaStream_t st);
void cb_knn_replay_launch(int K, int m, const float *xyz, const float *new_xyz, const int *offset,
                          const int *new_offset, int b, int *idx, float *dist2, int sqrt_dist, const CbGridView &v,
                          cudaStream_t st);
void cb_knn_reset_flagged(const CbGridView &v, cudaStream_t st);
// radius flavour of the query (TF batch_ordered_neighbors): nanoflann distance arithmetic, entries with
// d2 >= r2 replaced by pad_idx, ties left in search order (the reference's std::sort leaves them unspecified)
int cb_knn_query_radius_impl(int m, int K, const float *xyz, int n, const float *new_xyz, const int *offset,
                             const int *new_offset, int b, int *idx, float r2, int pad_idx, const CbGridView &v,
                             cudaStream_t st);
__global__ void k_bbox_init(unsigned *bbox, int *occ, int b, CbGridHeader *hdr, int n);
__global__ void k_bbox(const float *__restrict__ xyz, int n, const int *__restrict__ offset, int b, unsigned *bbox);

// ---------------------------------------------------------------------------------------------
// Warp-resident sorted top-K list: entry e = j*32 + lane lives in register j of lane `lane`.
// ---------------------------------------------------------------------------------------------
template <int KPL>
struct CbTopK {
    float d[KPL];
    int i[KPL];
    float kth;      // current K-th smallest squared distance (uniform across the warp)
    float tie_val;  // d2 at which a candidate was dropped/evicted with equality to the then-kth
    int K, lane;

    __device__ __forceinline__ void init(int K_, int lane_, int pad_idx)
    {
        K = K_; lane = lane_;
#pragma unroll
        for (int j = 0; j < KPL; j++) { d[j] = 1e10f; i[j] = pad_idx; }   // knnquery_cuda_kernel.cu:91-94
        kth = 1e10f;
        tie_val = -1.f;
    }

    __device__ __forceinline__ float entry_d(int e) const
    {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < KPL; j++) {
            float t = __shfl_sync(CB_FULL_MASK, d[j], e & 31);
            if ((e >> 5) == j) v = t;
        }
        return v;
    }

    // insert candidate (cd, ci), uniform across the warp, known to satisfy cd < kth
    __device__ __forceinline__ void insert(float cd, int ci)
    {
        int p = 0;
#pragma unroll
        for (int j = 0; j < KPL; j++) p += __popc(__ballot_sync(CB_FULL_MASK, d[j] <= cd));
        const float old_kth = kth;
#pragma unroll
        for (int j = KPL - 1; j >= 0; j--) {
            float up_d = __shfl_up_sync(CB_FULL_MASK, d[j], 1);
            int up_i = __shfl_up_sync(CB_FULL_MASK, i[j], 1);
            if (j > 0) {
                float w_d = __shfl_sync(CB_FULL_MASK, d[j - 1], 31);
                int w_i = __shfl_sync(CB_FULL_MASK, i[j - 1], 31);
                if (lane == 0) { up_d = w_d; up_i = w_i; }
            }
            const int e = j * 32 + lane;
            if (e > p) { d[j] = up_d; i[j] = up_i; }
            else if (e == p) { d[j] = cd; i[j] = ci; }
        }
        kth = entry_d(K - 1);
        if (kth == old_kth) tie_val = kth;   // the evicted K-th equals the new K-th: boundary tie
    }

    // process one candidate per lane (valid lanes only)
    __device__ __forceinline__ void offer(bool valid, float cd, int ci)
    {
        unsigned eq = __ballot_sync(CB_FULL_MASK, valid && cd == kth);
        if (eq) tie_val = kth;
        unsigned pass = __ballot_sync(CB_FULL_MASK, valid && cd < kth);
        while (pass) {
            const int src = __ffs(pass) - 1;
            pass &= pass - 1;
            const float sd = __shfl_sync(CB_FULL_MASK, cd, src);
            const int si = __shfl_sync(CB_FULL_MASK, ci, src);
            if (sd < kth) insert(sd, si);
            else if (sd == kth) tie_val = kth;
        }
    }

    __device__ __forceinline__ float out_d(int j) const { return d[j]; }
    __device__ __forceinline__ int out_i(int j) const { return i[j]; }

    // true if the SET of the K nearest is ambiguous: a candidate outside the list ties with the K-th entry
    __device__ __forceinline__ bool has_boundary_tie() const { return (tie_val == kth) && (kth < 1e10f); }
    // true if the reference's result for this query may depend on its heap mechanics
    __device__ __forceinline__ bool has_tie() const
    {
        bool t = (tie_val == kth) && (kth < 1e10f);
#pragma unroll
        for (int j = 0; j < KPL; j++) {
            float nx_d = __shfl_down_sync(CB_FULL_MASK, d[j], 1);
            if (j + 1 < KPL) {
                float w = __shfl_sync(CB_FULL_MASK, d[j + 1], 0);
                if (lane == 31) nx_d = w;
            }
            const int e = j * 32 + lane;
            const bool last_reg_last_lane = (j + 1 == KPL) && lane == 31;
            bool adj = !last_reg_last_lane && (e + 1 < K) && (d[j] == nx_d) && (d[j] < 1e10f);
            t = t || __any_sync(CB_FULL_MASK, adj);
        }
        return t;
    }
};

// ---------------------------------------------------------------------------------------------
// Bitonic variant for K <= 64: the list holds the 32*KPL smallest (d2, idx) keys seen so far as
// 64-bit keys (d2 >= 0, so its float bits order like an unsigned int).  A batch of 32 candidates is
// pre-filtered against the (K+1)-th smallest, and the survivors are either inserted one by one (few)
// or warp-sorted with a 15-stage bitonic network and merged with the list (many): ~130 instructions
// per batch however many candidates enter, instead of ~25 per inserted candidate.
// ---------------------------------------------------------------------------------------------
typedef unsigned long long cb_key;
#define CB_KEY_INF 0xffffffffffffffffull

__device__ __forceinline__ cb_key cb_make_key(float d, int i) { return ((cb_key)__float_as_uint(d) << 32) | (unsigned)i; }
__device__ __forceinline__ float cb_key_d(cb_key k) { return __uint_as_float((unsigned)(k >> 32)); }
__device__ __forceinline__ int cb_key_i(cb_key k) { return (int)(unsigned)(k & 0xffffffffu); }
__device__ __forceinline__ cb_key cb_shfl_key(cb_key k, int src)
{
    return ((cb_key)__shfl_sync(CB_FULL_MASK, (unsigned)(k >> 32), src) << 32) | __shfl_sync(CB_FULL_MASK, (unsigned)k, src);
}
__device__ __forceinline__ cb_key cb_shfl_xor_key(cb_key k, int m)
{
    return ((cb_key)__shfl_xor_sync(CB_FULL_MASK, (unsigned)(k >> 32), m) << 32) | __shfl_xor_sync(CB_FULL_MASK, (unsigned)k, m);
}
// Compare-exchange networks on (d bits, idx) pairs.  Only the distance is compared; on equal distances
// both lanes keep their own element (ties are re-done by the exact replay anyway), which keeps the
// multiset intact with a single 32-bit compare per stage.
// `flip` is 0 for a lane that keeps the minimum of the pair and 0xffffffff for one that keeps the maximum
// (a > b  <=>  ~a < ~b for unsigned), so one compare serves both directions.
__device__ __forceinline__ void cb_cmpx(unsigned &d, unsigned &i, int j, unsigned flip)
{
    const unsigned pd = __shfl_xor_sync(CB_FULL_MASK, d, j), pi = __shfl_xor_sync(CB_FULL_MASK, i, j);
    const bool take = (pd ^ flip) < (d ^ flip);
    d = take ? pd : d;
    i = take ? pi : i;
}
// per-lane direction masks of the 15-stage sort (bit s) and the 5-stage merge, computed once per thread
__device__ __forceinline__ unsigned cb_sort_dirs(int lane)
{
    unsigned m = 0;
    int s = 0;
    for (int size = 2; size <= 32; size <<= 1) {
        const bool up = size == 32 ? true : ((lane & size) == 0);
        for (int j = size >> 1; j > 0; j >>= 1, s++)
            if ((((lane & j) == 0) == up) == false) m |= 1u << s;       // bit set -> keep max
    }
    return m;
}
// ascending bitonic merge of a bitonic 32-sequence held one key per lane
__device__ __forceinline__ cb_key cb_bitonic_merge32(cb_key k, int lane)
{
    unsigned d = (unsigned)(k >> 32), i = (unsigned)k;
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) cb_cmpx(d, i, j, (lane & j) ? 0xffffffffu : 0u);
    return ((cb_key)d << 32) | i;
}
// full ascending bitonic sort of 32 keys
__device__ __forceinline__ cb_key cb_bitonic_sort32(cb_key k, unsigned dirs)
{
    unsigned d = (unsigned)(k >> 32), i = (unsigned)k;
    int s = 0;
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int j = size >> 1; j > 0; j >>= 1, s++) cb_cmpx(d, i, j, 0u - ((dirs >> s) & 1u));
    }
    return ((cb_key)d << 32) | i;
}

template <int KPL>
struct CbTopKB {
    cb_key key[KPL];    // entry e = j*32 + lane, ascending
    float kth;          // K-th smallest d2 (entry K-1), uniform
    float thr;          // pre-filter threshold: (K+1)-th smallest (entry min(K, 32*KPL-1))
    float tie_val;      // only needed when K == 32*KPL: a dropped / evicted value equal to the then-last entry
    int K, lane, cap_e;
    unsigned dirs;

    __device__ __forceinline__ void init(int K_, int lane_, int pad_idx)
    {
        K = K_; lane = lane_;
        dirs = cb_sort_dirs(lane_);
        cap_e = K < 32 * KPL - 1 ? K : 32 * KPL - 1;
#pragma unroll
        for (int j = 0; j < KPL; j++) key[j] = cb_make_key(1e10f, pad_idx);
        kth = 1e10f; thr = 1e10f; tie_val = -1.f;
    }
    __device__ __forceinline__ cb_key entry(int e) const
    {
        cb_key v = 0;
#pragma unroll
        for (int j = 0; j < KPL; j++) {
            const cb_key t = cb_shfl_key(key[j], e & 31);
            if ((e >> 5) == j) v = t;
        }
        return v;
    }
    __device__ __forceinline__ void refresh_thresholds()
    {
        kth = cb_key_d(entry(K - 1));
        thr = cb_key_d(entry(cap_e));
    }
    // sorted insertion of one key (uniform across the warp), shifting the tail right by one
    __device__ __forceinline__ void insert(cb_key c)
    {
        int p = 0;
#pragma unroll
        for (int j = 0; j < KPL; j++) p += __popc(__ballot_sync(CB_FULL_MASK, (unsigned)(key[j] >> 32) <= (unsigned)(c >> 32)));
#pragma unroll
        for (int j = KPL - 1; j >= 0; j--) {
            cb_key up = ((cb_key)__shfl_up_sync(CB_FULL_MASK, (unsigned)(key[j] >> 32), 1) << 32) |
                        __shfl_up_sync(CB_FULL_MASK, (unsigned)key[j], 1);
            if (j > 0) {
                const cb_key w = cb_shfl_key(key[j - 1], 31);
                if (lane == 0) up = w;
            }
            const int e = j * 32 + lane;
            if (e > p) key[j] = up;
            else if (e == p) key[j] = c;
        }
    }
    __device__ __forceinline__ void offer(bool valid, float cd, int ci)
    {
        if (K == 32 * KPL) {   // no spare entry: remember candidates dropped with equality to the last entry
            if (__ballot_sync(CB_FULL_MASK, valid && cd == thr)) tie_val = thr;
        }
        unsigned pass = __ballot_sync(CB_FULL_MASK, valid && cd < thr);
        if (!pass) return;
        cb_key c = (valid && cd < thr) ? cb_make_key(cd, ci) : CB_KEY_INF;
        if (__popc(pass) <= 5) {
            while (pass) {
                const int src = __ffs(pass) - 1;
                pass &= pass - 1;
                if (K == 32 * KPL) {        // the evicted entry is the current last one
                    // `pass` was taken against the threshold at the START of the batch: after an earlier insertion of
                    // this batch the candidate may no longer belong to the list.  Only a candidate that really evicts
                    // (or equals) the last entry can create a boundary tie; one above it is simply dropped.
                    const cb_key cc = cb_shfl_key(c, src);
                    const float cdv = cb_key_d(cc), old_last = cb_key_d(entry(32 * KPL - 1));
                    if (cdv < old_last) {
                        insert(cc);
                        if (cb_key_d(entry(32 * KPL - 1)) == old_last) tie_val = old_last;
                    } else if (cdv == old_last) {
                        tie_val = old_last;
                    }
                } else {
                    insert(cb_shfl_key(c, src));
                }
            }
        } else {
            c = cb_bitonic_sort32(c, dirs);
            // merge into register 0, carry the upper half into the next register
#pragma unroll
            for (int j = 0; j < KPL; j++) {
                const cb_key r = cb_shfl_key(c, 31 - lane);
                const bool r_less = (unsigned)(r >> 32) < (unsigned)(key[j] >> 32);
                const cb_key lo = r_less ? r : key[j], hi = r_less ? key[j] : r;
                key[j] = cb_bitonic_merge32(lo, lane);
                if (j + 1 < KPL) {
                    c = cb_bitonic_merge32(hi, lane);
                } else if (K == 32 * KPL) {
                    // evicted = hi; a boundary tie exists iff its smallest REAL value equals the new last entry
                    const unsigned hb = (unsigned)(hi >> 32);
                    const unsigned mn = __reduce_min_sync(CB_FULL_MASK, hb);
                    const float new_last = cb_key_d(cb_shfl_key(key[j], 31));
                    if (__uint_as_float(mn) == new_last) tie_val = new_last;
                }
            }
        }
        refresh_thresholds();
    }
    // the SET of the K nearest is ambiguous: full list -> a dropped / evicted value equal to the last entry;
    // list with a spare entry -> entries K-1 and K (the best candidate outside the set) tie
    __device__ __forceinline__ bool has_boundary_tie() const
    {
        if (K == 32 * KPL) return (tie_val == thr) && (thr < 1e10f);
        return (kth == thr) && (kth < 1e10f);
    }
    __device__ __forceinline__ bool has_tie() const
    {
        bool t = (K == 32 * KPL) && (tie_val == thr) && (thr < 1e10f);
#pragma unroll
        for (int j = 0; j < KPL; j++) {
            cb_key nx = ((cb_key)__shfl_down_sync(CB_FULL_MASK, (unsigned)(key[j] >> 32), 1) << 32) |
                        __shfl_down_sync(CB_FULL_MASK, (unsigned)key[j], 1);
            if (j + 1 < KPL) {
                const cb_key w = cb_shfl_key(key[j + 1], 0);
                if (lane == 31) nx = w;
            }
            const int e = j * 32 + lane;
            const bool last = (j + 1 == KPL) && lane == 31;
            const float d0 = cb_key_d(key[j]), d1 = cb_key_d(nx);
            const bool adj = !last && (e + 1 <= cap_e) && (d0 == d1) && (d0 < 1e10f);
            t = t || __any_sync(CB_FULL_MASK, adj);
        }
        return t;
    }
    __device__ __forceinline__ float out_d(int j) const { return cb_key_d(key[j]); }
    __device__ __forceinline__ int out_i(int j) const { return cb_key_i(key[j]); }
};

// Per-warp scratch for flattening cell-row ranges into candidate slots.
struct CbWarpScratch {
    int start[32];
    int excl[33];
};

template <int KPL> struct CbTopKSel { typedef CbTopK<KPL> type; };
template <> struct CbTopKSel<1> { typedef CbTopKB<1> type; };
template <> struct CbTopKSel<2> { typedef CbTopKB<2> type; };

// Search the grid for the K nearest supports of query (qx,qy,qz) in scene `sc`.
// Returns false if the query must be replayed by the exact brute-force kernel instead
// (too many rings).  All lanes of the warp call this together.
// squared distance in the arithmetic of the operator being replaced:
//   mode 0  pointops CUDA kernels (fma contraction, common.cuh cb_sqdist)
//   mode 1  nanoflann L2_Simple_Adaptor on the host (no fma): ((0 + dx*dx) + dy*dy) + dz*dz, d = q - p
//           (tensorflow/ops/tf_custom_ops/cpp_utils/nanoflann/nanoflann.hpp:432-440)
__device__ __forceinline__ float cb_sqdist_mode(int mode, float qx, float qy, float qz, float px, float py, float pz)
{
    if (mode == 0) return cb_sqdist(qx, qy, qz, px, py, pz);
    const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

template <class TK>
__device__ __forceinline__ bool cb_grid_search(TK &tk, const CbScene &sc, float qx, float qy, float qz,
                                               const int *__restrict__ cells, const float4 *__restrict__ sorted,
                                               CbWarpScratch *ws, int lane, int dist_mode = 0)
{
    const float fx = fminf(fmaxf(cb_cellf(qx, sc.ox, sc.inv_h), -1.0e6f), 1.0e6f);
    const float fy = fminf(fmaxf(cb_cellf(qy, sc.oy, sc.inv_h), -1.0e6f), 1.0e6f);
    const float fz = fminf(fmaxf(cb_cellf(qz, sc.oz, sc.inv_h), -1.0e6f), 1.0e6f);
    const int cx = (int)floorf(fx), cy = (int)floorf(fy), cz = (int)floorf(fz);
    const float rx = fx - (float)cx, ry = fy - (float)cy, rz = fz - (float)cz;
    int r0 = 0;
    r0 = max(r0, max(-cx, cx - (sc.nx - 1)));
    r0 = max(r0, max(-cy, cy - (sc.ny - 1)));
    r0 = max(r0, max(-cz, cz - (sc.nz - 1)));
    if (r0 > CB_KNN_MAX_RING) return false;

    const int r_first = max(r0, 1);     // first pass = the whole (2r+1)^3 block (cells nearer than r0 lie outside the grid)
    for (int r = r_first;; r++) {
        // rows (y,z) of shell r that fall inside the grid
        const int ya = max(cy - r, 0), yb = min(cy + r, sc.ny - 1);
        const int za = max(cz - r, 0), zb = min(cz + r, sc.nz - 1);
        const int wy = yb - ya + 1, wz = zb - za + 1;
        const int nrows = (wy > 0 && wz > 0) ? wy * wz : 0;
        for (int rb = 0; rb < nrows; rb += 32) {
            const int row = rb + lane;
            int s0 = 0, l0 = 0, s1 = 0, l1 = 0;
            if (row < nrows) {
                const int y = ya + row % wy, z = za + row / wy;
                const int dy = y - cy, dz = z - cz;
                const int rowbase = sc.cell_base + (z * sc.ny + y) * sc.nx;
                const bool outer = (r == r_first) || (abs(dy) == r) || (abs(dz) == r);
                if (outer) {
                    const int xa = max(cx - r, 0), xb = min(cx + r, sc.nx - 1);
                    if (xa <= xb) {
                        s0 = __ldg(cells + rowbase + xa);
                        l0 = __ldg(cells + rowbase + xb + 1) - s0;
                    }
                } else {
                    const int xl = cx - r, xh = cx + r;
                    if (xl >= 0 && xl < sc.nx) {
                        s0 = __ldg(cells + rowbase + xl);
                        l0 = __ldg(cells + rowbase + xl + 1) - s0;
                    }
                    if (xh >= 0 && xh < sc.nx) {
                        s1 = __ldg(cells + rowbase + xh);
                        l1 = __ldg(cells + rowbase + xh + 1) - s1;
                    }
                }
            }
#pragma unroll 1
            for (int seg = 0; seg < 2; seg++) {
                const int st = seg ? s1 : s0, ln = seg ? l1 : l0;
                if (seg == 1 && !__any_sync(CB_FULL_MASK, ln > 0)) break;
                // inclusive scan of lengths
                int inc = ln;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int t = __shfl_up_sync(CB_FULL_MASK, inc, o);
                    if (lane >= o) inc += t;
                }
                const int total = __shfl_sync(CB_FULL_MASK, inc, 31);
                if (total == 0) continue;
                __syncwarp();
                ws->start[lane] = st;
                ws->excl[lane] = inc - ln;
                if (lane == 31) ws->excl[32] = total;
                __syncwarp();
                for (int t = 0; t < total; t += 32) {
                    const int j = t + lane;
                    const bool valid = j < total;
                    float cd = 0.f;
                    int ci = 0;
                    if (valid) {
                        // largest l with excl[l] <= j
                        int lo = 0, hi = 31;
#pragma unroll
                        for (int it = 0; it < 5; it++) {
                            const int mid = (lo + hi + 1) >> 1;
                            if (ws->excl[mid] <= j) lo = mid; else hi = mid - 1;
                        }
                        const float4 c = __ldg(sorted + ws->start[lo] + (j - ws->excl[lo]));
                        cd = cb_sqdist_mode(dist_mode, qx, qy, qz, c.x, c.y, c.z);
                        ci = __float_as_int(c.w);
                    }
                    tk.offer(valid, cd, ci);
                }
            }
        }
        // coverage after shell r: distance (in cells) from the query to the nearest face of the
        // searched block beyond which unseen cells exist
        float bound = 3.0e38f;
        if (cx - r > 0) bound = fminf(bound, rx + (float)r);
        if (cx + r < sc.nx - 1) bound = fminf(bound, (float)(r + 1) - rx);
        if (cy - r > 0) bound = fminf(bound, ry + (float)r);
        if (cy + r < sc.ny - 1) bound = fminf(bound, (float)(r + 1) - ry);
        if (cz - r > 0) bound = fminf(bound, rz + (float)r);
        if (cz + r < sc.nz - 1) bound = fminf(bound, (float)(r + 1) - rz);
        if (bound > 1.0e38f) return true;                 // whole grid covered
        const float bd = (bound - 0.02f) * sc.h;          // 0.02 cell safety for fp32 cell assignment
        if (bd > 0.f && tk.kth < bd * bd * 0.9999f) return true;
        if (r + 1 > CB_KNN_MAX_RING) return false;
    }
}
