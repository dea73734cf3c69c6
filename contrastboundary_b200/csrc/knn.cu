// knn.cu — a1: grid-hash K-nearest-neighbour query, bit-identical to the reference's brute-force
// heap kernel (pytorch/lib/pointops/src/knnquery/knnquery_cuda_kernel.cu:65-119).
//
// Pipeline (all on the caller's stream, no host sync, no allocation):
//   bbox -> trial grid (volume heuristic) -> occupancy + local-dimension estimate -> final cell
//   size -> counting sort of the supports by cell (x-fastest cell order, so a row of cells is a
//   contiguous point range) -> warp-per-query ring search with a register-resident sorted top-K
//   -> exact replay of the reference heap for the few queries whose answer depends on it (ties).
#include "knn.cuh"
#include <stdarg.h>
#include <string.h>

// ---------------------------------------------------------------------------------------------
// error string
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void cb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char *cb_last_error_string(void) { return g_err; }
extern "C" int cb_version(void) { return 100; }
unsigned long long g_cb_launches = 0;
int g_cb_pdl = 1;          // programmatic dependent launch for the kernels that support it (common.cuh)
extern "C" int cb_set_pdl(int on)
{
    if (on == 0 || on == 1) g_cb_pdl = on;
    return g_cb_pdl;
}
extern "C" unsigned long long cb_launch_count(void) { return g_cb_launches; }

// ---------------------------------------------------------------------------------------------
// workspace layout
// ---------------------------------------------------------------------------------------------
static int cb_cell_cap(int n, int b) { return 16 * n + 4096 * b + 4096; }
static int cb_trial_cap(int n, int b) { return 2 * n + 512 * b + 512; }

size_t cb_grid_layout(int n, int m, int b, void *base, CbGridView *v)
{
    size_t off = 0;
    char *p = (char *)base;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = cb_align_up(off + bytes, 256);
        return p ? (void *)(p + o) : (void *)nullptr;
    };
    CbGridView t;
    t.cell_cap = cb_cell_cap(n, b);
    t.trial_cap = cb_trial_cap(n, b);
    t.max_tiles = (t.cell_cap + 1 + CB_SCAN_TILE - 1) / CB_SCAN_TILE;
    t.hdr = (CbGridHeader *)take(sizeof(CbGridHeader));
    t.scenes = (CbScene *)take(sizeof(CbScene) * (size_t)b);
    t.bbox = (unsigned *)take(sizeof(unsigned) * 6 * (size_t)b);
    t.occ = (int *)take(sizeof(int) * 2 * (size_t)b);
    t.tile_sums = (int *)take(sizeof(int) * (size_t)t.max_tiles);
    t.cells = (int *)take(sizeof(int) * ((size_t)t.cell_cap + 1));
    t.coarse = (int *)take(sizeof(int) * ((size_t)t.trial_cap + 1));
    t.point_cell = (int *)take(sizeof(int) * (size_t)n);
    t.point_rank = (int *)take(sizeof(int) * (size_t)n);
    t.sorted = (float4 *)take(sizeof(float4) * (size_t)n);
    t.flagged = (int *)take(sizeof(int) * (size_t)(m > 0 ? m : 1));
    if (v) *v = t;
    return off;
}

extern "C" size_t cb_knn_workspace_bytes(int n, int m, int b)
{
    if (n < 0 || m < 0 || b <= 0) return 0;
    return cb_grid_layout(n, m, b, nullptr, nullptr);
}

// ---------------------------------------------------------------------------------------------
// grid build kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_bbox_init(unsigned *bbox, int *occ, int b, CbGridHeader *hdr, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b * 6) bbox[i] = (i % 6 < 3) ? 0xffffffffu : 0u;   // min slots / max slots
    if (i < b * 2) occ[i] = 0;
    if (i == 0) { hdr->total_cells = 0; hdr->flagged_count = 0; hdr->n = n; hdr->b = b; hdr->trial = 1; }
}

__global__ void k_bbox(const float *__restrict__ xyz, int n, const int *__restrict__ offset, int b, unsigned *bbox)
{
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + lane;
        const bool valid = i < n;
        const int ic = valid ? i : n - 1;
        const int s = cb_scene_of(ic, offset, b);
        const unsigned ux = cb_f2ord(__ldg(xyz + 3 * ic)), uy = cb_f2ord(__ldg(xyz + 3 * ic + 1)),
                       uz = cb_f2ord(__ldg(xyz + 3 * ic + 2));
        const int s0 = __shfl_sync(CB_FULL_MASK, s, 0);
        if (__all_sync(CB_FULL_MASK, s == s0)) {
            unsigned mnx = __reduce_min_sync(CB_FULL_MASK, ux), mny = __reduce_min_sync(CB_FULL_MASK, uy),
                     mnz = __reduce_min_sync(CB_FULL_MASK, uz);
            unsigned mxx = __reduce_max_sync(CB_FULL_MASK, ux), mxy = __reduce_max_sync(CB_FULL_MASK, uy),
                     mxz = __reduce_max_sync(CB_FULL_MASK, uz);
            if (lane == 0) {
                unsigned *bb = bbox + 6 * s0;
                atomicMin(bb + 0, mnx); atomicMin(bb + 1, mny); atomicMin(bb + 2, mnz);
                atomicMax(bb + 3, mxx); atomicMax(bb + 4, mxy); atomicMax(bb + 5, mxz);
            }
        } else if (valid) {
            unsigned *bb = bbox + 6 * s;
            atomicMin(bb + 0, ux); atomicMin(bb + 1, uy); atomicMin(bb + 2, uz);
            atomicMax(bb + 3, ux); atomicMax(bb + 4, uy); atomicMax(bb + 5, uz);
        }
    }
}

static float g_occ_factor = 0.45f;   // points per occupied cell = factor * K (tuning knob, cb_knn_set_occupancy)
extern "C" float cb_knn_set_occupancy(float f)
{
    if (f > 0.05f && f < 4.f) g_occ_factor = f;
    return g_occ_factor;
}
__device__ __forceinline__ float cb_target_occ(int nsample, float factor)
{
    // points per occupied cell that makes the 3x3x3 block hold the K nearest for most queries
    return fminf(fmaxf(factor * (float)nsample, 2.0f), 48.0f);
}

// dims for cell size h; returns number of cells
__device__ __forceinline__ long long cb_dims(float ex, float ey, float ez, float h, int *nx, int *ny, int *nz)
{
    const float inv = 1.0f / h;
    *nx = (int)fminf(floorf(ex * inv), (float)(CB_GRID_MAX_DIM - 1)) + 1;
    *ny = (int)fminf(floorf(ey * inv), (float)(CB_GRID_MAX_DIM - 1)) + 1;
    *nz = (int)fminf(floorf(ez * inv), (float)(CB_GRID_MAX_DIM - 1)) + 1;
    return (long long)*nx * *ny * *nz;
}

// phase 0: trial grid from the bbox-volume heuristic; phase 1: final grid from measured occupancy
__device__ void cb_params_block(CbGridHeader *hdr, CbScene *scenes, const unsigned *bbox, const int *occ,
                                const int *__restrict__ offset, int b, int n, int nsample, int cap, int phase, float occ_factor)
{
    const float target = cb_target_occ(nsample, occ_factor);
    for (int s = threadIdx.x; s < b; s += blockDim.x) {
        CbScene sc;
        sc.start = s == 0 ? 0 : offset[s - 1];
        sc.end = offset[s];
        sc.pad = 0;
        const int ns = sc.end - sc.start;
        if (ns <= 0) {
            sc.ox = sc.oy = sc.oz = 0.f; sc.h = 1.f; sc.inv_h = 1.f; sc.nx = sc.ny = sc.nz = 1; sc.cell_base = 0;
            scenes[s] = sc;
            continue;
        }
        const unsigned *bb = bbox + 6 * s;
        sc.ox = cb_ord2f(bb[0]); sc.oy = cb_ord2f(bb[1]); sc.oz = cb_ord2f(bb[2]);
        float ex = cb_ord2f(bb[3]) - sc.ox, ey = cb_ord2f(bb[4]) - sc.oy, ez = cb_ord2f(bb[5]) - sc.oz;
        if (!(ex >= 0.f) || !(ey >= 0.f) || !(ez >= 0.f) || ex > 1e30f || ey > 1e30f || ez > 1e30f) { ex = ey = ez = 0.f; }
        const float emax = fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-20f));
        float h;
        if (phase == 0) {
            const float vx = fmaxf(ex, 1e-3f * emax), vy = fmaxf(ey, 1e-3f * emax), vz = fmaxf(ez, 1e-3f * emax);
            h = cbrtf(target * vx * vy * vz / (float)ns);
        } else {
            const CbScene old = scenes[s];
            const float m1 = (float)max(occ[2 * s], 1), m2 = (float)max(occ[2 * s + 1], 1);
            float dim = log2f(fmaxf(m1 / m2, 1.0f));
            dim = fminf(fmaxf(dim, 1.5f), 3.0f);
            const float o0 = (float)ns / m1;
            h = old.h * powf(target / o0, 1.0f / dim);
        }
        h = fmaxf(h, emax / (float)(CB_GRID_MAX_DIM - 1));
        h = fmaxf(h, 1e-20f);
        // per-scene share of the cell budget
        const long long budget = max((long long)((double)(cap - 64 * b - 1) * (double)ns / (double)max(n, 1)), 64LL);
        int nx, ny, nz;
        for (int it = 0; it < 400; it++) {
            if (cb_dims(ex, ey, ez, h, &nx, &ny, &nz) <= budget) break;
            h *= 1.1f;
        }
        cb_dims(ex, ey, ez, h, &nx, &ny, &nz);
        sc.h = h; sc.inv_h = 1.0f / h; sc.nx = nx; sc.ny = ny; sc.nz = nz; sc.cell_base = 0;
        scenes[s] = sc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int base = 0;
        for (int s = 0; s < b; s++) {
            scenes[s].cell_base = base;
            const int ns = scenes[s].end - scenes[s].start;
            if (ns > 0) base += scenes[s].nx * scenes[s].ny * scenes[s].nz;
        }
        hdr->total_cells = base;
        hdr->trial = phase == 0;
    }
}

__global__ void k_params(CbGridHeader *hdr, CbScene *scenes, const unsigned *bbox, const int *occ,
                         const int *__restrict__ offset, int b, int n, int nsample, int cap, int phase, float occ_factor)
{
    cb_params_block(hdr, scenes, bbox, occ, offset, b, n, nsample, cap, phase, occ_factor);
}

// count points per cell.  trial: also measure occupied cells at h and 2h.
__global__ void k_count(const float *__restrict__ xyz, int n, const int *__restrict__ offset, int b,
                        const CbScene *__restrict__ scenes, int *cells, int *coarse, int *occ, int *point_cell,
                        int *point_rank, int trial)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int s = cb_scene_of(i, offset, b);
        const CbScene sc = scenes[s];
        const float x = __ldg(xyz + 3 * i), y = __ldg(xyz + 3 * i + 1), z = __ldg(xyz + 3 * i + 2);
        int cx = (int)floorf(cb_cellf(x, sc.ox, sc.inv_h)), cy = (int)floorf(cb_cellf(y, sc.oy, sc.inv_h)),
            cz = (int)floorf(cb_cellf(z, sc.oz, sc.inv_h));
        cx = min(max(cx, 0), sc.nx - 1); cy = min(max(cy, 0), sc.ny - 1); cz = min(max(cz, 0), sc.nz - 1);
        const int cell = sc.cell_base + (cz * sc.ny + cy) * sc.nx + cx;
        const int rank = atomicAdd(cells + cell, 1);
        if (trial) {
            if (rank == 0) atomicAdd(occ + 2 * s, 1);
            const int hx = (sc.nx + 1) >> 1, hy = (sc.ny + 1) >> 1;
            const int cc = sc.cell_base + ((cz >> 1) * hy + (cy >> 1)) * hx + (cx >> 1);
            if (atomicExch(coarse + cc, 1) == 0) atomicAdd(occ + 2 * s + 1, 1);
        } else {
            point_cell[i] = cell;
            point_rank[i] = rank;
        }
    }
}

// exclusive scan over cells[0 .. total_cells] (inclusive of the sentinel), in place, two kernels
__global__ void __launch_bounds__(256) k_scan_tiles(int *cells, int *tile_sums, const CbGridHeader *hdr)
{
    const int total = hdr->total_cells + 1;
    const int tile0 = blockIdx.x * CB_SCAN_TILE;
    if (tile0 >= total) return;
    __shared__ int warp_sums[8];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    int v[8];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int i = tile0 + t * 8 + k;
        v[k] = i < total ? cells[i] : 0;
        sum += v[k];
    }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(CB_FULL_MASK, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    int woff = 0;
    for (int k = 0; k < w; k++) woff += warp_sums[k];
    int run = woff + inc - sum;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int i = tile0 + t * 8 + k;
        if (i < total) cells[i] = run;
        run += v[k];
    }
    if (t == 255) tile_sums[blockIdx.x] = run;
}

__global__ void __launch_bounds__(256) k_scan_add(int *cells, const int *tile_sums, const CbGridHeader *hdr)
{
    const int total = hdr->total_cells + 1;
    const int tile0 = blockIdx.x * CB_SCAN_TILE;
    if (tile0 >= total || blockIdx.x == 0) return;
    __shared__ int red[256];
    int acc = 0;
    for (int k = threadIdx.x; k < blockIdx.x; k += 256) acc += tile_sums[k];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    const int off = red[0];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int i = tile0 + threadIdx.x * 8 + k;
        if (i < total) cells[i] += off;
    }
}

__global__ void k_fill(const float *__restrict__ xyz, int n, const int *__restrict__ cells,
                       const int *__restrict__ point_cell, const int *__restrict__ point_rank, float4 *sorted)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int pos = cells[point_cell[i]] + point_rank[i];
        sorted[pos] = make_float4(__ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2), __int_as_float(i));
    }
}

// zero cells[0 .. total_cells] and the trial's coarse flags (sizes are device-known)
__global__ void k_zero_cells(int *cells, int *coarse, const CbGridHeader *hdr, int zero_coarse)
{
    const int total = hdr->total_cells + 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        cells[i] = 0;
        if (zero_coarse) coarse[i] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// The whole grid build as ONE cooperative kernel: the ten phases above separated by grid-wide barriers instead of
// kernel boundaries (a build of a 40960-point scene is ~6 us of work behind ~55 us of launch latency otherwise).
// Same arithmetic, same results as the multi-kernel path (the counting sort's ranks are atomics in both).
// ---------------------------------------------------------------------------------------------
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(256) k_grid_build_fused(const float *__restrict__ xyz, int n, const int *__restrict__ offset, int b,
                                                          int nsample, float occ_factor, CbGridHeader *hdr, CbScene *scenes,
                                                          unsigned *bbox, int *occ, int *tile_sums, int *cells, int *coarse,
                                                          int *point_cell, int *point_rank, float4 *sorted, int trial_cap,
                                                          int cell_cap)
{
    cg::grid_group grid = cg::this_grid();
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    // phase 0: init
    for (int i = gtid; i < b * 6; i += gsize) bbox[i] = (i % 6 < 3) ? 0xffffffffu : 0u;
    for (int i = gtid; i < b * 2; i += gsize) occ[i] = 0;
    if (gtid == 0) { hdr->total_cells = 0; hdr->flagged_count = 0; hdr->n = n; hdr->b = b; hdr->trial = 1; }
    grid.sync();
    // phase 1: bounding boxes
    {
        const int lane = threadIdx.x & 31;
        for (int base = gtid - lane; base < n; base += gsize) {
            const int i = base + lane;
            const bool valid = i < n;
            const int ic = valid ? i : n - 1;
            const int s = cb_scene_of(ic, offset, b);
            const unsigned ux = cb_f2ord(__ldg(xyz + 3 * ic)), uy = cb_f2ord(__ldg(xyz + 3 * ic + 1)), uz = cb_f2ord(__ldg(xyz + 3 * ic + 2));
            const int s0 = __shfl_sync(CB_FULL_MASK, s, 0);
            if (__all_sync(CB_FULL_MASK, s == s0)) {
                unsigned mnx = __reduce_min_sync(CB_FULL_MASK, ux), mny = __reduce_min_sync(CB_FULL_MASK, uy), mnz = __reduce_min_sync(CB_FULL_MASK, uz);
                unsigned mxx = __reduce_max_sync(CB_FULL_MASK, ux), mxy = __reduce_max_sync(CB_FULL_MASK, uy), mxz = __reduce_max_sync(CB_FULL_MASK, uz);
                if (lane == 0) {
                    unsigned *bb = bbox + 6 * s0;
                    atomicMin(bb + 0, mnx); atomicMin(bb + 1, mny); atomicMin(bb + 2, mnz);
                    atomicMax(bb + 3, mxx); atomicMax(bb + 4, mxy); atomicMax(bb + 5, mxz);
                }
            } else if (valid) {
                unsigned *bb = bbox + 6 * s;
                atomicMin(bb + 0, ux); atomicMin(bb + 1, uy); atomicMin(bb + 2, uz);
                atomicMax(bb + 3, ux); atomicMax(bb + 4, uy); atomicMax(bb + 5, uz);
            }
        }
    }
    grid.sync();
    for (int phase = 0; phase < 2; phase++) {
        // trial grid (phase 0) / final grid (phase 1): parameters, zero the cell counters, count
        if (blockIdx.x == 0) cb_params_block(hdr, scenes, bbox, occ, offset, b, n, nsample, phase == 0 ? trial_cap : cell_cap, phase, occ_factor);
        grid.sync();
        const int total = hdr->total_cells + 1;
        for (int i = gtid; i < total; i += gsize) {
            cells[i] = 0;
            if (phase == 0) coarse[i] = 0;
        }
        grid.sync();
        for (int i = gtid; i < n; i += gsize) {
            const int s = cb_scene_of(i, offset, b);
            const CbScene sc = scenes[s];
            const float x = __ldg(xyz + 3 * i), y = __ldg(xyz + 3 * i + 1), z = __ldg(xyz + 3 * i + 2);
            int cx = (int)floorf(cb_cellf(x, sc.ox, sc.inv_h)), cy = (int)floorf(cb_cellf(y, sc.oy, sc.inv_h)),
                cz = (int)floorf(cb_cellf(z, sc.oz, sc.inv_h));
            cx = min(max(cx, 0), sc.nx - 1); cy = min(max(cy, 0), sc.ny - 1); cz = min(max(cz, 0), sc.nz - 1);
            const int cell = sc.cell_base + (cz * sc.ny + cy) * sc.nx + cx;
            const int rank = atomicAdd(cells + cell, 1);
            if (phase == 0) {
                if (rank == 0) atomicAdd(occ + 2 * s, 1);
                const int hx = (sc.nx + 1) >> 1, hy = (sc.ny + 1) >> 1;
                const int cc = sc.cell_base + ((cz >> 1) * hy + (cy >> 1)) * hx + (cx >> 1);
                if (atomicExch(coarse + cc, 1) == 0) atomicAdd(occ + 2 * s + 1, 1);
            } else {
                point_cell[i] = cell;
                point_rank[i] = rank;
            }
        }
        grid.sync();
    }
    // exclusive scan of cells[0 .. total_cells]: per-tile scans, then the tile offsets
    const int total = hdr->total_cells + 1;
    const int ntiles = (total + CB_SCAN_TILE - 1) / CB_SCAN_TILE;
    __shared__ int warp_sums[8];
    __shared__ int red[256];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tile0 = tile * CB_SCAN_TILE;
        int v[8];
        int sum = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = tile0 + t * 8 + k;
            v[k] = i < total ? cells[i] : 0;
            sum += v[k];
        }
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(CB_FULL_MASK, inc, o);
            if (lane >= o) inc += u;
        }
        __syncthreads();                       // warp_sums of the previous tile are consumed
        if (lane == 31) warp_sums[w] = inc;
        __syncthreads();
        int woff = 0;
        for (int k = 0; k < w; k++) woff += warp_sums[k];
        int run = woff + inc - sum;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = tile0 + t * 8 + k;
            if (i < total) cells[i] = run;
            run += v[k];
        }
        if (t == 255) tile_sums[tile] = run;
    }
    grid.sync();
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (tile == 0) continue;
        int acc = 0;
        for (int k = t; k < tile; k += 256) acc += tile_sums[k];
        __syncthreads();                       // red[] of the previous tile is consumed
        red[t] = acc;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (t < s) red[t] += red[t + s];
            __syncthreads();
        }
        const int off = red[0];
        const int tile0 = tile * CB_SCAN_TILE;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = tile0 + t * 8 + k;
            if (i < total) cells[i] += off;
        }
    }
    grid.sync();
    // counting-sort scatter: supports re-ordered by cell as (x, y, z, original index)
    for (int i = gtid; i < n; i += gsize) {
        const int pos = cells[point_cell[i]] + point_rank[i];
        sorted[pos] = make_float4(__ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2), __int_as_float(i));
    }
}

static int g_grid_fused = 1;   // 1: one cooperative kernel (default) | 0: the multi-kernel build | 2: fused unless the stream is capturing
extern "C" int cb_grid_set_fused(int mode)
{
    if (mode >= 0 && mode <= 2) g_grid_fused = mode;
    return g_grid_fused;
}

// ---------------------------------------------------------------------------------------------
// query kernel: one warp per query
// ---------------------------------------------------------------------------------------------
template <int KPL>
__global__ void __launch_bounds__(128) k_knn_query(int m, int K, const float *__restrict__ new_xyz,
                                                   const int *__restrict__ new_offset, int b, int self_query,
                                                   const CbScene *__restrict__ scenes, const int *__restrict__ cells,
                                                   const float4 *__restrict__ sorted, int *__restrict__ idx,
                                                   float *__restrict__ dist2, int sqrt_dist, CbGridHeader *hdr,
                                                   int *flagged, int dist_mode, float r2, int pad_idx)
{
    __shared__ CbWarpScratch scratch[4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int w = blockIdx.x * 4 + wib;
    if (w >= m) return;
    int q = w;
    float qx, qy, qz;
    if (self_query) {   // walk queries in cell order: neighbouring warps touch the same cells
        const float4 p = __ldg(sorted + w);
        q = __float_as_int(p.w); qx = p.x; qy = p.y; qz = p.z;
    } else {
        qx = __ldg(new_xyz + 3 * q); qy = __ldg(new_xyz + 3 * q + 1); qz = __ldg(new_xyz + 3 * q + 2);
    }
    const int s = cb_scene_of(q, new_offset, b);
    const CbScene sc = scenes[s];
    typename CbTopKSel<KPL>::type tk;
    tk.init(K, lane, sc.start);
    bool ok = cb_grid_search(tk, sc, qx, qy, qz, cells, sorted, &scratch[wib], lane, dist_mode);
    // sqrt_dist bit 1 = "set semantics": the caller only uses the SET of neighbours (label histograms), so ties INSIDE the set
    // need no replay of the reference heap; only a tie across the K-th boundary does
    if (ok && dist_mode == 0 && ((sqrt_dist & 2) ? tk.has_boundary_tie() : tk.has_tie())) ok = false;
    if (!ok) {
        if (lane == 0) flagged[atomicAdd(&hdr->flagged_count, 1)] = q;
        return;
    }
#pragma unroll
    for (int j = 0; j < KPL; j++) {
        const int e = j * 32 + lane;
        if (e < K) {
            if (dist_mode == 0) {
                idx[(size_t)q * K + e] = tk.out_i(j);
                dist2[(size_t)q * K + e] = (sqrt_dist & 1) ? __fsqrt_rn(tk.out_d(j)) : tk.out_d(j);
            } else {
                idx[(size_t)q * K + e] = tk.out_d(j) < r2 ? tk.out_i(j) : pad_idx;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// exact replay of the reference kernel for flagged queries: one block per flagged query, the
// block evaluates distances in index order, thread 0 runs the reference's heap
// (knnquery_cuda_kernel.cu:21-48,91-110) in shared memory.
// ---------------------------------------------------------------------------------------------
// The reference's reheap (knnquery_cuda_kernel.cu:21-36) with the value on its way down kept in registers: place (vd, vi)
// at the root of the max-heap and sift it down; returns the new root distance.  Same comparisons in the same order as the
// reference's swap loop (dist[root] there IS the moving value), so the same arrangement results.  The heap is stored as
// packed (distance, index) pairs, node j at hp[j + 1]: the two children of a node are one aligned 16-byte shared-memory
// load, and while the larger child is being chosen the child pairs of BOTH children are already on their way (the
// storage holds 2 * K + 8 entries so that these speculative loads stay inside it) — the shared-memory latency of a level
// overlaps with the compare-and-branch of the level above.  This loop is the serial critical path of a replay.
// (explicit shared-space accesses on a 32-bit address: the generic-pointer form re-derives the shared window from
// SR_CgaCtaId inside the loop)
__device__ __forceinline__ int4 cb_lds128(unsigned a)
{
    int4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ int2 cb_lds64(unsigned a)
{
    int2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void cb_sts64(unsigned a, int x, int y)
{
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ float cb_replay_sift(unsigned hp_s, int len, float vd, int vi)
{
    int r = 0, child = 1;
    float new_root = vd;
    int4 c = cb_lds128(hp_s + 16);                                          // nodes 1, 2
    while (child < len) {
        const int4 na = cb_lds128(hp_s + 16 * child + 16);                  // children of node child
        const int4 nb = cb_lds128(hp_s + 16 * child + 32);                  // children of node child + 1
        const bool right = child + 1 < len && __int_as_float(c.z) > __int_as_float(c.x);
        const float cdv = __int_as_float(right ? c.z : c.x);
        const int cvi = right ? c.w : c.y;
        if (vd > cdv) break;
        cb_sts64(hp_s + 8 * r + 8, __float_as_int(cdv), cvi);
        if (r == 0) new_root = cdv;
        r = child + (right ? 1 : 0);
        child = r * 2 + 1;
        c.x = right ? nb.x : na.x; c.y = right ? nb.y : na.y; c.z = right ? nb.z : na.z; c.w = right ? nb.w : na.w;
    }
    cb_sts64(hp_s + 8 * r + 8, __float_as_int(vd), vi);
    return new_root;
}

#define CB_REPLAY_THREADS 1024
#define CB_REPLAY_BATCH 4096                  /* capacity of the survivor list of one batch */
#define CB_REPLAY_ROWS 8                      /* rows of 32 candidates a warp keeps in flight */
static size_t cb_replay_smem(int K);
// The heap replay itself is serial (the reference's result under ties is defined by its heap
// history), but only candidates with d2 < root are ever inserted and the root never grows.  So the
// block filters a batch of candidates in parallel against the root at the start of the batch
// (an upper bound for every later root), compacts the survivors IN INDEX ORDER, and thread 0
// replays just those.  Batches DOUBLE (256, 512, ... ): a batch of s candidates after p scanned ones
// leaves about s * K / p survivors, so s = p keeps the serial pass at ~K entries per batch while the
// number of block-wide synchronisation rounds is log2(n / 256) instead of n / 4096.  Warp w owns a contiguous
// segment of the batch and walks it in coalesced rows of 32, CB_REPLAY_ROWS rows of loads in flight:
// (warp, row, lane) order is index order.  Pass 1 counts survivors, pass 2 (same arithmetic, L1/L2 hits) stores them
// behind the scanned offsets; if a batch leaves more than CB_REPLAY_BATCH survivors (an adversarial point order), it
// is retried at a quarter of the size, down to CB_REPLAY_BATCH candidates — which always fit.
__global__ void __launch_bounds__(CB_REPLAY_THREADS) k_knn_replay(int K, const float *__restrict__ xyz,
                                                                 const float *__restrict__ new_xyz,
                                                                 const int *__restrict__ offset,
                                                                 const int *__restrict__ new_offset, int b,
                                                                 int *__restrict__ idx, float *__restrict__ dist2,
                                                                 int sqrt_dist, const CbGridHeader *hdr,
                                                                 const int *__restrict__ flagged, int dist_mode, float r2,
                                                                 int pad_idx)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int2 *hp = (int2 *)smem_raw;                                           // 2K + 8 packed heap entries, node j at hp[j + 1]
    int2 *cand = hp + 2 * K + 8;                                           // CB_REPLAY_BATCH survivors (distance, index)
    __shared__ int warp_cnt[CB_REPLAY_THREADS / 32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int count = hdr->flagged_count;
    for (int f = blockIdx.x; f < count; f += gridDim.x) {
        const int q = flagged[f];
        const int s = cb_scene_of(q, new_offset, b);
        const int start = s == 0 ? 0 : offset[s - 1], end = offset[s];
        const float qx = new_xyz[3 * q], qy = new_xyz[3 * q + 1], qz = new_xyz[3 * q + 2];
        for (int k = t; k < 2 * K + 8; k += CB_REPLAY_THREADS) hp[k] = make_int2(__float_as_int(1e10f), start);
        __syncthreads();
        int bs = 256;
        for (int base = start; base < end;) {
            const float root = __int_as_float(hp[1].x);
            const int cur = min(bs, end - base);
            const int rows_total = (cur + 31) >> 5;
            const int rows_per_warp = (rows_total + CB_REPLAY_THREADS / 32 - 1) / (CB_REPLAY_THREADS / 32);
            const int r0 = w * rows_per_warp, r1 = min(rows_total, r0 + rows_per_warp);
            const int lim = base + cur;
            int cnt = 0;                                                    // pass 1: survivors of this warp's segment
            for (int r = r0; r < r1; r += CB_REPLAY_ROWS) {
                float px[CB_REPLAY_ROWS], py[CB_REPLAY_ROWS], pz[CB_REPLAY_ROWS];
#pragma unroll
                for (int u = 0; u < CB_REPLAY_ROWS; u++) {
                    const int i = base + ((r + u) << 5) + lane;
                    const bool in = r + u < r1 && i < lim;
                    px[u] = in ? __ldg(xyz + 3 * i) : 0.f;
                    py[u] = in ? __ldg(xyz + 3 * i + 1) : 0.f;
                    pz[u] = in ? __ldg(xyz + 3 * i + 2) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < CB_REPLAY_ROWS; u++) {
                    const int i = base + ((r + u) << 5) + lane;
                    const bool in = r + u < r1 && i < lim;
                    const float d = in ? cb_sqdist_mode(dist_mode, qx, qy, qz, px[u], py[u], pz[u]) : 3.0e38f;
                    cnt += __popc(__ballot_sync(CB_FULL_MASK, d < root));
                }
            }
            if (lane == 0) warp_cnt[w] = cnt;
            __syncthreads();
            int woff, total;                                                // exclusive scan over the warps' counts
            {
                const int wc = lane < CB_REPLAY_THREADS / 32 ? warp_cnt[lane] : 0;
                int winc = wc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int v = __shfl_up_sync(CB_FULL_MASK, winc, o);
                    if (lane >= o) winc += v;
                }
                woff = __shfl_sync(CB_FULL_MASK, winc - wc, w);
                total = __shfl_sync(CB_FULL_MASK, winc, 31);
            }
            if (total > CB_REPLAY_BATCH) {                                  // does not fit: retry this batch smaller
                bs = max(bs >> 2, CB_REPLAY_BATCH);
                __syncthreads();
                continue;
            }
            if (cnt > 0) {                                                  // pass 2: store in (warp, row, lane) = index order
                int pos = woff;
                for (int r = r0; r < r1; r += CB_REPLAY_ROWS) {
                    float px[CB_REPLAY_ROWS], py[CB_REPLAY_ROWS], pz[CB_REPLAY_ROWS];
#pragma unroll
                    for (int u = 0; u < CB_REPLAY_ROWS; u++) {
                        const int i = base + ((r + u) << 5) + lane;
                        const bool in = r + u < r1 && i < lim;
                        px[u] = in ? __ldg(xyz + 3 * i) : 0.f;
                        py[u] = in ? __ldg(xyz + 3 * i + 1) : 0.f;
                        pz[u] = in ? __ldg(xyz + 3 * i + 2) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < CB_REPLAY_ROWS; u++) {
                        const int i = base + ((r + u) << 5) + lane;
                        const bool in = r + u < r1 && i < lim;
                        const float d = in ? cb_sqdist_mode(dist_mode, qx, qy, qz, px[u], py[u], pz[u]) : 3.0e38f;
                        const unsigned bal = __ballot_sync(CB_FULL_MASK, d < root);
                        if (d < root) cand[pos + __popc(bal & lt_mask)] = make_int2(__float_as_int(d), i);
                        pos += __popc(bal);
                    }
                }
            }
            __syncthreads();
            if (t == 0 && total > 0) {
                const unsigned hp_s = (unsigned)__cvta_generic_to_shared(hp), cand_s = (unsigned)__cvta_generic_to_shared(cand);
                float rootv = __int_as_float(cb_lds64(hp_s + 8).x);
                int2 e = cb_lds64(cand_s);
                for (int u = 0; u < total; u++) {
                    const int2 nxt = cb_lds64(cand_s + 8 * (u + 1 < total ? u + 1 : u));   // next survivor on its way during the sift
                    if (__int_as_float(e.x) < rootv) rootv = cb_replay_sift(hp_s, K, __int_as_float(e.x), e.y);
                    e = nxt;
                }
            }
            __syncthreads();
            base += cur;
            if (bs < (1 << 28)) bs <<= 1;
        }
        if (t == 0) {                                                       // heap_sort (:39-48)
            for (int i = K - 1; i > 0; i--) {
                const unsigned hp_s = (unsigned)__cvta_generic_to_shared(hp);
                const int2 v = cb_lds64(hp_s + 8 * i + 8), top = cb_lds64(hp_s + 8);
                cb_sts64(hp_s + 8 * i + 8, top.x, top.y);
                (void)cb_replay_sift(hp_s, i, __int_as_float(v.x), v.y);
            }
        }
        __syncthreads();
        for (int k = t; k < K; k += CB_REPLAY_THREADS) {
            const int2 e = hp[k + 1];
            const float dk = __int_as_float(e.x);
            if (dist_mode == 0) {
                idx[(size_t)q * K + k] = e.y;
                dist2[(size_t)q * K + k] = (sqrt_dist & 1) ? __fsqrt_rn(dk) : dk;
            } else {
                idx[(size_t)q * K + k] = dk < r2 ? e.y : pad_idx;
            }
        }
        __syncthreads();
    }
}

static size_t cb_replay_smem(int K)
{
    static bool opted_in = false;                  // rows wider than ~2040 entries need more than the default 48 KB
    if (!opted_in) {
        cudaFuncSetAttribute(k_knn_replay, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        opted_in = true;
    }
    return (2 * (size_t)K + 8) * 8 + (size_t)CB_REPLAY_BATCH * 8;
}

// brute-force for every query (K > 256): flag all, then replay
__global__ void k_flag_all(int m, CbGridHeader *hdr, int *flagged)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) flagged[i] = i;
    if (i == 0) hdr->flagged_count = m;
}

__global__ void k_reset_flagged(CbGridHeader *hdr) { hdr->flagged_count = 0; }

// ---------------------------------------------------------------------------------------------
// host entry points
// ---------------------------------------------------------------------------------------------
static int grid_blocks(int n, int threads) { int g = (n + threads - 1) / threads; return g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g); }

static int cb_grid_build_fused(const float *xyz, int n, const int *offset, int b, int nsample_hint, const CbGridView &v,
                               cudaStream_t st)
{
    static int max_blocks = 0;
    if (max_blocks == 0) {
        int per_sm = 0, dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_grid_build_fused, 256, 0);
        max_blocks = per_sm * sms;
        if (max_blocks < 1) max_blocks = -1;
    }
    if (max_blocks < 0) return 1;
    int blocks = (n + 255) / 256;
    if (blocks > 148 * 3) blocks = 148 * 3;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    float occ = g_occ_factor;
    CbGridHeader *hdr = v.hdr; CbScene *scenes = v.scenes; unsigned *bbox = v.bbox; int *occp = v.occ, *tile_sums = v.tile_sums;
    int *cells = v.cells, *coarse = v.coarse, *point_cell = v.point_cell, *point_rank = v.point_rank;
    float4 *sorted = v.sorted;
    int trial_cap = v.trial_cap, cell_cap = v.cell_cap;
    void *args[] = {(void *)&xyz, (void *)&n, (void *)&offset, (void *)&b, (void *)&nsample_hint, (void *)&occ, (void *)&hdr,
                    (void *)&scenes, (void *)&bbox, (void *)&occp, (void *)&tile_sums, (void *)&cells, (void *)&coarse,
                    (void *)&point_cell, (void *)&point_rank, (void *)&sorted, (void *)&trial_cap, (void *)&cell_cap};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_grid_build_fused, dim3(blocks), dim3(256), args, 0, st);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return 1;
    }
    CB_COUNT(1);
    return 0;
}

int cb_grid_build_impl(const float *xyz, int n, const int *offset, int b, int nsample_hint, const CbGridView &v,
                           cudaStream_t st)
{
    if (g_grid_fused && n > 0) {
        bool use = true;
        if (g_grid_fused == 2) {
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) use = false;
        }
        if (use && cb_grid_build_fused(xyz, n, offset, b, nsample_hint, v, st) == 0) {
            CB_CUDA_CHECK("cb_grid_build");
            return CB_OK;
        }
    }
    const int ib = (b * 6 + 127) / 128;
    k_bbox_init<<<ib, 128, 0, st>>>(v.bbox, v.occ, b, v.hdr, n);
    if (n > 0) k_bbox<<<grid_blocks(n, 256), 256, 0, st>>>(xyz, n, offset, b, v.bbox);
    // trial grid
    k_params<<<1, 128, 0, st>>>(v.hdr, v.scenes, v.bbox, v.occ, offset, b, n, nsample_hint, v.trial_cap, 0, g_occ_factor);
    k_zero_cells<<<148, 256, 0, st>>>(v.cells, v.coarse, v.hdr, 1);
    if (n > 0) k_count<<<grid_blocks(n, 256), 256, 0, st>>>(xyz, n, offset, b, v.scenes, v.cells, v.coarse, v.occ,
                                                             v.point_cell, v.point_rank, 1);
    // final grid
    k_params<<<1, 128, 0, st>>>(v.hdr, v.scenes, v.bbox, v.occ, offset, b, n, nsample_hint, v.cell_cap, 1, g_occ_factor);
    k_zero_cells<<<148 * 2, 256, 0, st>>>(v.cells, v.coarse, v.hdr, 0);
    if (n > 0) k_count<<<grid_blocks(n, 256), 256, 0, st>>>(xyz, n, offset, b, v.scenes, v.cells, v.coarse, v.occ,
                                                             v.point_cell, v.point_rank, 0);
    k_scan_tiles<<<v.max_tiles, 256, 0, st>>>(v.cells, v.tile_sums, v.hdr);
    k_scan_add<<<v.max_tiles, 256, 0, st>>>(v.cells, v.tile_sums, v.hdr);
    if (n > 0) k_fill<<<grid_blocks(n, 256), 256, 0, st>>>(xyz, n, v.cells, v.point_cell, v.point_rank, v.sorted);
    CB_COUNT(11);
    CB_CUDA_CHECK("cb_grid_build");
    return CB_OK;
}

static int check_common(int m, int nsample, const void *xyz, int n, const void *offset, const void *new_offset, int b)
{
    CB_REQUIRE(m >= 0 && n >= 0 && b > 0, CB_EINVAL, "cb_knn: bad sizes m=%d n=%d b=%d", m, n, b);
    CB_REQUIRE(nsample >= 1 && nsample <= CB_KNN_MAX_NSAMPLE, CB_EINVAL, "cb_knn: nsample=%d outside [1,%d]", nsample,
               CB_KNN_MAX_NSAMPLE);
    CB_REQUIRE((xyz || n == 0) && offset && new_offset, CB_EINVAL, "cb_knn: NULL pointer argument");
    return CB_OK;
}

extern "C" int cb_grid_build(const float *xyz, int n, const int *offset, int b, int nsample_hint, void *grid,
                             size_t grid_bytes, void *stream)
{
    CB_REQUIRE(n >= 0 && b > 0 && offset && grid && (xyz || n == 0), CB_EINVAL, "cb_grid_build: bad arguments");
    CbGridView v;
    const size_t need = cb_grid_layout(n, 0, b, grid, &v);
    CB_REQUIRE(grid_bytes >= need, CB_EWORKSPACE, "cb_grid_build: workspace %zu < %zu", grid_bytes, need);
    CB_REQUIRE(((uintptr_t)grid & 255) == 0, CB_EINVAL, "cb_grid_build: workspace not 256-byte aligned");
    if (nsample_hint < 1) nsample_hint = 16;
    return cb_grid_build_impl(xyz, n, offset, b, nsample_hint, v, (cudaStream_t)stream);
}

void cb_knn_replay_launch(int K, int m, const float *xyz, const float *new_xyz, const int *offset,
                          const int *new_offset, int b, int *idx, float *dist2, int sqrt_dist, const CbGridView &v,
                          cudaStream_t st)
{
    const size_t smem = cb_replay_smem(K);
    const int rblocks = K <= 256 ? 148 : (m < 148 * 8 ? (m > 0 ? m : 1) : 148 * 8);
    k_knn_replay<<<rblocks, CB_REPLAY_THREADS, smem, st>>>(K, xyz, new_xyz, offset, new_offset, b, idx, dist2,
                                                           sqrt_dist, v.hdr, v.flagged, 0, 0.f, 0);
}

int cb_knn_query_radius_impl(int m, int K, const float *xyz, int n, const float *new_xyz, const int *offset,
                             const int *new_offset, int b, int *idx, float r2, int pad_idx, const CbGridView &v,
                             cudaStream_t st)
{
    if (m == 0 || K == 0) return CB_OK;
    CB_REQUIRE(K <= 2048, CB_EUNSUPPORTED, "radius neighbours: row width %d > 2048 unsupported", K);
    const int self_query = (new_xyz == xyz && m == n) ? 1 : 0;
    const int blocks = (m + 3) / 4;
    const size_t smem = cb_replay_smem(K);
    if (K > 256) {
        // rows wider than the register top-K of k_knn_query (dense clouds during neighbourhood-limit calibration,
        // datasets/base.py:199-294): every query takes the exact scene scan
        k_flag_all<<<(m + 255) / 256, 256, 0, st>>>(m, v.hdr, v.flagged);
        const int rblocks = m < 148 * 8 ? m : 148 * 8;
        k_knn_replay<<<rblocks, CB_REPLAY_THREADS, smem, st>>>(K, xyz, new_xyz, offset, new_offset, b, idx, nullptr, 0,
                                                               v.hdr, v.flagged, 1, r2, pad_idx);
        CB_COUNT(2);
        CB_CUDA_CHECK("cb_batch_radius_neighbors");
        return CB_OK;
    }
    k_reset_flagged<<<1, 1, 0, st>>>(v.hdr);
#define CB_LAUNCH_R(KPL)                                                                                          \
    k_knn_query<KPL><<<blocks, 128, 0, st>>>(m, K, new_xyz, new_offset, b, self_query, v.scenes, v.cells, v.sorted, \
                                              idx, nullptr, 0, v.hdr, v.flagged, 1, r2, pad_idx)
    if (K <= 32) CB_LAUNCH_R(1);
    else if (K <= 64) CB_LAUNCH_R(2);
    else if (K <= 128) CB_LAUNCH_R(4);
    else CB_LAUNCH_R(8);
#undef CB_LAUNCH_R
    k_knn_replay<<<148, CB_REPLAY_THREADS, smem, st>>>(K, xyz, new_xyz, offset, new_offset, b, idx, nullptr, 0, v.hdr,
                                                       v.flagged, 1, r2, pad_idx);
    CB_COUNT(3);
    CB_CUDA_CHECK("cb_batch_radius_neighbors");
    return CB_OK;
}

void cb_knn_reset_flagged(const CbGridView &v, cudaStream_t st) { k_reset_flagged<<<1, 1, 0, st>>>(v.hdr); }

static int query_impl(int m, int K, const float *xyz, int n, const float *new_xyz, const int *offset,
                      const int *new_offset, int b, int *idx, float *dist2, int sqrt_dist, const CbGridView &v,
                      cudaStream_t st)
{
    if (m == 0) return CB_OK;
    const int self_query = (new_xyz == xyz && m == n) ? 1 : 0;
    const int blocks = (m + 3) / 4;
    if (K <= 256) {
        k_reset_flagged<<<1, 1, 0, st>>>(v.hdr);
#define CB_LAUNCH_Q(KPL)                                                                                          \
    k_knn_query<KPL><<<blocks, 128, 0, st>>>(m, K, new_xyz, new_offset, b, self_query, v.scenes, v.cells, v.sorted, \
                                              idx, dist2, sqrt_dist, v.hdr, v.flagged, 0, 0.f, 0)
        if (K <= 32) CB_LAUNCH_Q(1);
        else if (K <= 64) CB_LAUNCH_Q(2);
        else if (K <= 128) CB_LAUNCH_Q(4);
        else CB_LAUNCH_Q(8);
#undef CB_LAUNCH_Q
    } else {
        k_flag_all<<<(m + 255) / 256, 256, 0, st>>>(m, v.hdr, v.flagged);
    }
    cb_knn_replay_launch(K, m, xyz, new_xyz, offset, new_offset, b, idx, dist2, sqrt_dist, v, st);
    CB_COUNT(3);
    CB_CUDA_CHECK("cb_knn_query");
    return CB_OK;
}

extern "C" int cb_knn_query_grid(int m, int nsample, const float *xyz, int n, const float *new_xyz, const int *offset,
                                 const int *new_offset, int b, int *idx, float *dist2, int sqrt_dist, void *grid,
                                 size_t grid_bytes, void *stream)
{
    int rc = check_common(m, nsample, xyz, n, offset, new_offset, b);
    if (rc) return rc;
    CB_REQUIRE(grid && (idx || m == 0) && (dist2 || m == 0), CB_EINVAL, "cb_knn_query_grid: NULL pointer argument");
    if (!new_xyz) new_xyz = xyz;
    CbGridView v;
    const size_t need = cb_grid_layout(n, m, b, grid, &v);
    CB_REQUIRE(grid_bytes >= need, CB_EWORKSPACE, "cb_knn_query_grid: workspace %zu < %zu", grid_bytes, need);
    return query_impl(m, nsample, xyz, n, new_xyz, offset, new_offset, b, idx, dist2, sqrt_dist, v, (cudaStream_t)stream);
}

extern "C" int cb_knn_query(int m, int nsample, const float *xyz, int n, const float *new_xyz, const int *offset,
                            const int *new_offset, int b, int *idx, float *dist2, int sqrt_dist, void *workspace,
                            size_t workspace_bytes, void *stream)
{
    int rc = check_common(m, nsample, xyz, n, offset, new_offset, b);
    if (rc) return rc;
    CB_REQUIRE(workspace && (idx || m == 0) && (dist2 || m == 0), CB_EINVAL, "cb_knn_query: NULL pointer argument");
    CB_REQUIRE(((uintptr_t)workspace & 255) == 0, CB_EINVAL, "cb_knn_query: workspace not 256-byte aligned");
    if (!new_xyz) new_xyz = xyz;
    CbGridView v;
    const size_t need = cb_grid_layout(n, m, b, workspace, &v);
    CB_REQUIRE(workspace_bytes >= need, CB_EWORKSPACE, "cb_knn_query: workspace %zu < %zu", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    if (nsample <= 256) {
        rc = cb_grid_build_impl(xyz, n, offset, b, nsample, v, st);
        if (rc) return rc;
    } else {
        k_bbox_init<<<1, 128, 0, st>>>(v.bbox, v.occ, 1, v.hdr, n);   // header only; brute force needs no grid
    }
    return query_impl(m, nsample, xyz, n, new_xyz, offset, new_offset, b, idx, dist2, sqrt_dist, v, st);
}
