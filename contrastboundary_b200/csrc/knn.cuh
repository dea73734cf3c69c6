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

// Per-warp scratch for flattening cell-row ranges into candidate slots.
struct CbWarpScratch {
    int start[32];
    int excl[33];
};

// Search the grid for the K nearest supports of query (qx,qy,qz) in scene `sc`.
// Returns false if the query must be replayed by the exact brute-force kernel instead
// (too many rings).  All lanes of the warp call this together.
template <int KPL>
__device__ __forceinline__ bool cb_grid_search(CbTopK<KPL> &tk, const CbScene &sc, float qx, float qy, float qz,
                                               const int *__restrict__ cells, const float4 *__restrict__ sorted,
                                               CbWarpScratch *ws, int lane)
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

    for (int r = r0;; r++) {
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
                const bool outer = (abs(dy) == r) || (abs(dz) == r);
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
                        cd = cb_sqdist(qx, qy, qz, c.x, c.y, c.z);
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
