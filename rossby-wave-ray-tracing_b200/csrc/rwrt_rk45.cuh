// rwrt_rk45.cuh -- the hot kernels: Dormand-Prince 5(4) with the reference's step controller,
// one ray per lane, lanes advance through output intervals independently and refill from a
// global queue when their ray finishes or dies (compaction of finished rays at warp granularity).
//
// This header is compiled twice (rwrt_rk45_fast.cu: FMA contraction on; rwrt_rk45_strict.cu:
// -fmad=false, every operation a separate IEEE op in the reference's order).  RWRT_VARIANT names
// the namespace so both sets of kernels live in one library.
//
// Reference semantics (file:line): rkf45.py:34-99 (initial step), 222-253 (step), 259-321
// (rk_step), 375-514 (_step_impl: clamp once per pass, retry loop, NaN-as-accept, controller),
// 601-615 (tableau); wr.py:767-887 (output-interval loop, pole / speed cut-offs, ug/vg at outputs).
#pragma once
#include "rwrt_device.cuh"
#include "rwrt_rk45_params.h"

#ifndef RWRT_VARIANT
#error "define RWRT_VARIANT (fast | strict)"
#endif

namespace rwrt {

namespace RWRT_VARIANT {

// rkf45.py:604-615
#define A21 (1.0 / 5)
#define A31 (3.0 / 40)
#define A32 (9.0 / 40)
#define A41 (44.0 / 45)
#define A42 (-56.0 / 15)
#define A43 (32.0 / 9)
#define A51 (19372.0 / 6561)
#define A52 (-25360.0 / 2187)
#define A53 (64448.0 / 6561)
#define A54 (-212.0 / 729)
#define A61 (9017.0 / 3168)
#define A62 (-355.0 / 33)
#define A63 (46732.0 / 5247)
#define A64 (49.0 / 176)
#define A65 (-5103.0 / 18656)
#define B1 (35.0 / 384)
#define B2 0.0
#define B3 (500.0 / 1113)
#define B4 (125.0 / 192)
#define B5 (-2187.0 / 6784)
#define B6 (11.0 / 84)
#define E1 (-71.0 / 57600)
#define E2 0.0
#define E3 (71.0 / 16695)
#define E4 (-71.0 / 1920)
#define E5 (17253.0 / 339200)
#define E6 (-22.0 / 525)
#define E7 (1.0 / 40)

__device__ __forceinline__ const double *member_base(const RunParams &p, long long slot)
{
    if (p.src_member == nullptr) return p.g.fpk;
    int isrc = (int)((slot / p.nzwn) % p.nsource);
    return p.g.fpk + (size_t)p.src_member[isrc] * p.g.member_stride;
}

// rkf45.py:29-31: squares summed component by component, sqrt, / sqrt(5)
__device__ __forceinline__ double rms5(double a0, double a1, double a2, double a3, double a4)
{
    double s = a0 * a0;
    s = s + a1 * a1;
    s = s + a2 * a2;
    s = s + a3 * a3;
    s = s + a4 * a4;
    return sqrt(s) / sqrt(5.0);
}

// the same with the division by sqrt(5) as a correctly rounded three-instruction sequence (same bits)
__device__ __forceinline__ double rms5_fast(double a0, double a1, double a2, double a3, double a4)
{
    double s = a0 * a0;
    s = s + a1 * a1;
    s = s + a2 * a2;
    s = s + a3 * a3;
    s = s + a4 * a4;
    return div_sqrt5(sqrt(s));
}

// rkf45.py:34-99 select_initial_step (direction +1, error-estimator order 4)
__device__ double initial_step(const Grid &g, const double *fb, const double *y0, const double *f0, double rtol,
                               double atol)
{
    double scale[5], y1[5], f1[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) scale[c] = atol + fabs(y0[c]) * rtol;
    double d0 = rms5(y0[0] / scale[0], y0[1] / scale[1], y0[2] / scale[2], y0[3] / scale[3], y0[4] / scale[4]);
    double d1 = rms5(f0[0] / scale[0], f0[1] / scale[1], f0[2] / scale[2], f0[3] / scale[3], f0[4] / scale[4]);
    double h0 = 0.01 * d0 / d1;
    if (d0 < 1e-5) h0 = 1e-6;
    if (d1 < 1e-5) h0 = 1e-6;
#pragma unroll
    for (int c = 0; c < 5; ++c) y1[c] = y0[c] + h0 * 1.0 * f0[c];
    ray_rhs<RWRT_FAST>(g, fb, y1, f1);
    double d2 = rms5((f1[0] - f0[0]) / scale[0], (f1[1] - f0[1]) / scale[1], (f1[2] - f0[2]) / scale[2],
                     (f1[3] - f0[3]) / scale[3], (f1[4] - f0[4]) / scale[4]) / h0;
    bool m1 = !(d1 > 1e-15), m2 = !(d2 > 1e-15);
    double mx = (d1 != d1) ? d2 : ((d2 != d2) ? d1 : fmax(d1, d2));  // np.nanmax
    double h1 = pow(0.01 / mx, 1.0 / (4 + 1));
    if (m1 && m2) h1 = np_maximum(1e-6, h0 * 1e-3);
    double hh = 100 * h0;
    if (hh != hh || h1 != h1) return qnan();  // np.min propagates NaN
    return hh < h1 ? hh : h1;
}

// Prepare: per live ray f0 = fun(y0) (rkf45.py:345) and the initial step (rkf45.py:347-349).
__global__ void __launch_bounds__(128) rk45_prepare(RunParams p)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.nlive) return;
    long long slot = p.live[i];
    const size_t pl = p.out_pl;
    double y[5], f[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) y[c] = p.out[c * pl + slot];
    const double *fb = member_base(p, slot);
    ray_rhs<RWRT_FAST>(p.g, fb, y, f);
    double h = initial_step(p.g, fb, y, f, p.rtol, p.atol);
    double *st = p.state + (size_t)i * RWRT_STATE;
#pragma unroll
    for (int c = 0; c < 5; ++c) { st[c] = y[c]; st[5 + c] = f[c]; }
    st[10] = h;
    st[11] = cos(y[1]);   // cos(latitude of the previous output)
    st[12] = 0.0;         // t
    st[13] = y[0];        // previous output position
    st[14] = y[1];
    reinterpret_cast<int *>(st + 15)[0] = 0;  // accepted so far
    reinterpret_cast<int *>(st + 15)[1] = 0;  // attempts so far; bit 31 set = ray finished
    if (p.rows_done) p.rows_done[i] = 1;      // next output row of this ray (the cooperative kernel's progress flag)
}

// Pre-fill of the output rows 1..nt-1, whole rows, fully coalesced (a streaming write at HBM speed):
//   * NaN-root slots: rlon, rlat, rzwn keep their t=0 value at every output time, the
//     other four planes are NaN; their step count is zero;
//   * live slots: NaN in all seven planes -- the integrator then overwrites the rows a ray is alive for, and a
//     ray that dies (wr.py:838-855) leaves its tail as it is instead of writing NaN rows one by one.
// Writing only the static slots (v1-v11) cost twice the time for a third of the bytes: static and live slots
// interleave below sector granularity, so DRAM did a read-modify-write for most sectors.
__global__ void __launch_bounds__(256) fill_static_slots(RunParams p)
{
    long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= p.N) return;
    const size_t pl = (size_t)p.nt * p.N;
    double v0 = p.out[0 * pl + slot], v1 = p.out[1 * pl + slot], v2 = p.out[2 * pl + slot], v3 = p.out[3 * pl + slot],
           v4 = p.out[4 * pl + slot];
    double s = v0 + v1 + v2 + v3 + v4;
    const double n = qnan();
    const bool live = (s == s);  // finite t=0 state: the integrator owns this slot
    if (live) v0 = v1 = v2 = v3 = v4 = n;
    // this block's share of the output rows (gridDim.y chunks); consecutive threads write consecutive slots
    int per = (p.nt - 1 + gridDim.y - 1) / gridDim.y;
    int it0 = 1 + blockIdx.y * per, it1 = min(p.nt, it0 + per);
    for (int it = it0; it < it1; ++it) {
        size_t o = (size_t)it * p.N + slot;
        // static slot: the t=0 values stay (NaN root stays NaN, a NaN position stays NaN): the reference never
        // advances it; ug, vg at outputs: fresh sample + extent form (wr.py:856-865) -- NaN because l or the state is NaN
        __stcs(p.out + 0 * pl + o, v0);
        __stcs(p.out + 1 * pl + o, v1);
        __stcs(p.out + 2 * pl + o, v2);
        __stcs(p.out + 3 * pl + o, v3);
        __stcs(p.out + 4 * pl + o, v4);
        __stcs(p.out + 5 * pl + o, n);
        __stcs(p.out + 6 * pl + o, n);
    }
    if (blockIdx.y == 0 && !live) {
        if (p.accepted) p.accepted[slot] = 0;
        if (p.attempts) p.attempts[slot] = 0;
        if (p.status) p.status[slot] = 0;
    }
}

// Record layout, second half: rows [it_begin, it_end) of the reference layout from the per-ray records -- a
// transpose through shared memory.  A block owns a tile of 32 consecutive slots x 16 rows: its warps read the
// records of a ray 512 contiguous bytes at a time (8 rows x 64 B), and write, per (plane, row), 32 consecutive
// slots = 256 contiguous bytes.  NaN-root slots get their constant rows, rows beyond a
// ray's last valid row are NaN (wr.py:838-855).  This pass replaces the pre-fill: every output byte is written once.
#ifndef RWRT_EX_ROWS
#define RWRT_EX_ROWS 16
#endif
#ifndef RWRT_EX_SLOTS
#define RWRT_EX_SLOTS 32
#endif
constexpr int EX_SLOTS = RWRT_EX_SLOTS, EX_ROWS = RWRT_EX_ROWS, EX_STRIDE = EX_ROWS * 8 + 1;   // doubles per slot in the tile (odd: no conflicts)
static_assert(EX_SLOTS % 32 == 0 && EX_ROWS % 8 == 0 && EX_SLOTS * EX_STRIDE * 8 <= 48 * 1024, "expand_records tile");
#ifndef RWRT_EX_TILES
#define RWRT_EX_TILES 6
#endif
// A block walks EX_TILES row tiles of its 32 slots: the records of tile i + 1 are requested (into registers) before
// tile i is written out, so reads and writes overlap inside a block and the slot -> ray -> rows_done lookups are paid
// once per block instead of once per tile (the one-tile-per-block form ran at 4.7 TB/s, latency-bound on that
// three-level load chain with six blocks per SM).  The constant rows of no-root slots are put into the tile once.
#ifndef RWRT_EX_MINBLOCKS
#define RWRT_EX_MINBLOCKS 4
#endif
constexpr int EX_ROWS_PER_BLOCK = EX_ROWS * RWRT_EX_TILES;
__global__ void __launch_bounds__(256, RWRT_EX_MINBLOCKS) expand_records(RunParams p)
{
    __shared__ double tile[EX_SLOTS * EX_STRIDE];
    constexpr int Q = EX_SLOTS / 8, H = EX_ROWS / 8;
    const long long slot0 = (long long)blockIdx.x * EX_SLOTS;
    const int block_row0 = p.it_begin + blockIdx.y * EX_ROWS_PER_BLOCK;
    const int block_row1 = min(p.it_end, block_row0 + EX_ROWS_PER_BLOCK);
    const size_t pl = p.out_pl;                                             // planes of the t = 0 row (constants of no-root slots)
    const size_t dpl = p.dst_pl;                                            // planes of the destination rows
    const double n = qnan();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rr = lane >> 2, c = lane & 3;                                // row within a pass of 8, 16-byte piece of the record
    // per-thread view of its Q slots: record base and last valid row (rd = 0: no record -- no-root slot or beyond N)
    const double2 *rec[Q];
    int rd[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int s = warp + 8 * q;
        const long long slot = slot0 + s;
        rec[q] = nullptr;
        rd[q] = 0;
        if (slot >= p.N) continue;
        const int li = p.inv[slot];
        if (li < 0) {   // NaN-root slot: planes 0..4 keep their t = 0 value, ug / vg are NaN -- the same in every row
            double a = (2 * c < 5) ? p.out[(size_t)(2 * c) * pl + slot] : n;
            double b = (2 * c + 1 < 5) ? p.out[(size_t)(2 * c + 1) * pl + slot] : n;
            double *t = tile + s * EX_STRIDE;
#pragma unroll
            for (int h = 0; h < H; ++h) {
                t[(rr + 8 * h) * 8 + 2 * c] = a;
                t[(rr + 8 * h) * 8 + 2 * c + 1] = b;
            }
            if (blockIdx.y == 0 && p.it_begin == 1 && lane == 0) {
                if (p.accepted) p.accepted[slot] = 0;
                if (p.attempts) p.attempts[slot] = 0;
                if (p.status) p.status[slot] = 0;
            }
        } else {
            rd[q] = min(p.rows_done[li], p.it_end);
            rec[q] = reinterpret_cast<const double2 *>(p.rec + ((size_t)li * p.rec_rows) * 8) + c;
        }
    }
    double2 v[Q][H];
    auto request = [&](int row0) {      // records of the tile that starts at row0 -> registers
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const int it = row0 + rr + 8 * h;
                v[q][h] = make_double2(n, n);
                if (it < rd[q]) v[q][h] = __ldcs(rec[q] + (size_t)(it - p.rec_it0) * 4);
            }
    };
    request(block_row0);
    for (int row0 = block_row0; row0 < block_row1; row0 += EX_ROWS) {
        // registers -> tile (live slots only: the constant rows of the others are there already)
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            if (rec[q]) {
                double *t = tile + (warp + 8 * q) * EX_STRIDE;
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    t[(rr + 8 * h) * 8 + 2 * c] = v[q][h].x;
                    t[(rr + 8 * h) * 8 + 2 * c + 1] = v[q][h].y;
                }
            }
        }
        __syncthreads();
        if (row0 + EX_ROWS < block_row1) request(row0 + EX_ROWS);         // in flight while this tile is written out
        // store phase: warp w writes rows w, w + 8, ... of all seven planes; per (plane, row) EX_SLOTS consecutive
        // slots = 512 contiguous bytes, in back-to-back 256-byte stores
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const int r = warp + 8 * j, it = row0 + r;
            if (it >= block_row1) break;
#pragma unroll
            for (int plane = 0; plane < 7; ++plane) {
#pragma unroll
                for (int half = 0; half < EX_SLOTS / 32; ++half) {
                    const long long slot = slot0 + 32 * half + lane;
                    if (slot < p.N)
                        __stcs(p.dst + (size_t)plane * dpl + (size_t)(it - p.dst_it0) * p.N + slot, tile[(32 * half + lane) * EX_STRIDE + r * 8 + plane]);
                }
            }
        }
        __syncthreads();                                                   // the tile is free again
    }
}

// Per-lane ray state kept in registers; the seven stage vectors K1..K7 live in shared memory
// (ks[stage][component][thread], conflict-free), which keeps the kernel at <= 128 registers and
// 16 resident warps per SM instead of 8.
struct Lane {
    double y[5];
    double t, h_abs, plon, plat;
    long long slot;
    int acc, att;
    int it;         // next output row to produce
    bool rejected;  // a rejection happened since the last accepted step (rkf45.py:462-469)
    bool fresh;     // next attempt is the first of a pass: apply the min-step clamp (rkf45.py:383-387)
};

constexpr int RK_THREADS = 128;
constexpr int RK_SMEM_BYTES = 7 * 5 * RK_THREADS * (int)sizeof(double);

__device__ __forceinline__ void write_row(const RunParams &p, long long slot, int it, double lon, double lat, double k,
                                          double l, double a, double ug, double vg)
{
    const size_t pl = (size_t)p.nt * p.N;
    size_t o = (size_t)it * p.N + slot;
    p.out[0 * pl + o] = lon;
    p.out[1 * pl + o] = lat;
    p.out[2 * pl + o] = k;
    p.out[3 * pl + o] = l;
    p.out[4 * pl + o] = a;
    p.out[5 * pl + o] = ug;
    p.out[6 * pl + o] = vg;
}

// rkf45.py:453-455 / 474-475: 0.9 * err^-0.2 clipped to [0.2, 10].  The clip makes pow() unnecessary
// outside (5e-6, 2000): there the clipped value does not depend on pow's rounding.
__device__ __forceinline__ double step_factor(double err, bool accept)
{
    // one evaluation site for the power (it is an out-of-line 166-instruction routine in the strict build)
    double f = 0.0;
    if (err >= 5.0e-6 && err <= 2000.0) f = 0.9 * (RWRT_FAST ? pow_m02(err) : pow(err, -0.2));
    if (accept) {
        if (err < 5.0e-6) return 10.0;  // 0.9 * err^-0.2 > 10.3 (also err == 0, rkf45.py:455)
        return f < 10.0 ? f : 10.0;
    }
    if (err > 2000.0) return 0.2;       // 0.9 * err^-0.2 < 0.197
    return f > 0.2 ? f : 0.2;
}

// Output-time group velocity when the RHS mask fired at an output point (|l| >= 100 there): the
// reference still samples and evaluates wn.py:266-294 without the mask (wr.py:856-865).  Rare; kept out
// of line so the hot loop holds a single instance of the sampler.
__device__ __noinline__ double2 output_ugvg_slow(Grid g, const double *fb, double lon, double lat, double k, double l)
{
    Merc m = sample_mercator<RWRT_FAST>(g, fb, lon, lat);
    double2 r;
    ugvg_extent(m.u, m.v, m.qx, m.qy, k, l, r.x, r.y);
    return r;
}

// Stage coefficients: rows 1..5 = A (rkf45.py:605-612), row 6 = B (rkf45.py:613) -- the sums
// sum_j K_j c_sj are accumulated in j order exactly like the reference's einsum.
__constant__ double cC[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
    {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};

// y + (sum_{j<s} K_j c_sj) * h for stage s with compile-time coefficient indices (constant-bank operands)
#define RWRT_COMBO(S, EXPR)                          \
    case S:                                          \
        _Pragma("unroll") for (int c = 0; c < 5; ++c) ys[c] = L.y[c] + (EXPR) * h; \
        break;

// The integrator.  Persistent grid; every lane owns one ray at a time.  The six stages run through ONE
// instance of the right-hand side (a loop, not six inlined copies): the kernel stays inside the
// instruction cache (v2, fully unrolled, spent 27 % of its warp time on instruction-fetch stalls).
__global__ void __launch_bounds__(RK_THREADS, 4) rk45_integrate(RunParams p)
{
    extern __shared__ double ks[];  // [7][5][RK_THREADS]
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x;
    const int lane_id = tid & 31;
#define KS(s, c) ks[((s) * 5 + (c)) * RK_THREADS + tid]
    Lane L;
    L.slot = -1;
    const double *fb = p.g.fpk;
    bool queue_empty = false;
    long long my_total = 0;

    for (;;) {
        __syncwarp();
        // ---- refill idle lanes from the queue (one atomic per warp) ----
        bool idle = (L.slot < 0);
        unsigned idle_mask = __ballot_sync(FULL, idle);
        if (idle_mask == FULL && queue_empty) break;
        if (idle_mask != 0 && !queue_empty) {
            int leader = __ffs(idle_mask) - 1;
            unsigned long long base = 0;
            if (lane_id == leader) base = atomicAdd(p.queue, (unsigned long long)__popc(idle_mask));
            base = __shfl_sync(FULL, base, leader);
            if (base >= (unsigned long long)p.nlive) queue_empty = true;
            if (idle) {
                unsigned long long mine = base + __popc(idle_mask & ((1u << lane_id) - 1));
                if (mine < (unsigned long long)p.nlive) {
                    const double *st = p.state + (size_t)mine * RWRT_STATE;
                    L.slot = p.live[mine];
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        L.y[c] = st[c];
                        KS(0, c) = st[5 + c];
                    }
                    L.h_abs = st[10];
                    L.t = 0.0;
                    L.plon = L.y[0];
                    L.plat = L.y[1];
                    L.acc = 0;
                    L.att = 0;
                    L.it = 1;
                    L.rejected = false;
                    L.fresh = true;
                    fb = member_base(p, L.slot);
                }
            }
        }
        if (L.slot < 0) continue;

        // ---- one attempt (rkf45.py:410-502) ----
        double tb = L.it * p.tstep;  // wr.py:798-801
        if (L.it == p.nt - 1 && tb > p.ttotal) tb = p.ttotal;
        if (L.fresh) {  // rkf45.py:383-387, once per pass
            if (!(L.h_abs >= p.gmin) && L.h_abs == L.h_abs) L.h_abs = p.gmin;
            L.rejected = false;
            L.fresh = false;
        }
        double t_new = L.t + L.h_abs;  // rkf45.py:426-435
        if (t_new - tb > 0) t_new = tb;
        const double h = t_new - L.t;
        ++L.att;
        double ys[5], kn[5], ug = 0, vg = 0;
        bool masked = false;
#pragma unroll 1
        for (int s = 1; s <= 6; ++s) {
            // rkf45.py:306-311 / 316: y + (sum_{j<s} K_j c_sj) * h, the sum accumulated in j order
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                double a = KS(0, c) * cC[s][0];
                for (int j = 1; j < s; ++j) a = a + KS(j, c) * cC[s][j];
                ys[c] = L.y[c] + a * h;
            }
            ray_rhs<RWRT_FAST>(p.g, fb, ys, kn, ug, vg, masked);
#pragma unroll
            for (int c = 0; c < 5; ++c) KS(s, c) = kn[c];
        }
        // now ys = y_new and KS(6, .) = kn = f(y_new) (rkf45.py:317-319)
        double r[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) {  // rkf45.py:442-446
            double scale = p.atol + np_maximum(fabs(L.y[c]), fabs(ys[c])) * p.rtol;
            double e = KS(0, c) * E1 + KS(1, c) * E2 + KS(2, c) * E3 + KS(3, c) * E4 + KS(4, c) * E5 + KS(5, c) * E6 +
                       kn[c] * E7;
            r[c] = (h * e) / scale;
        }
        double err = rms5(r[0], r[1], r[2], r[3], r[4]);
        if (err != err) err = 0;
        const bool accept = err < 1;
        double factor = step_factor(err, accept);
        if (accept) {
            if (L.rejected && factor > 1.0) factor = 1.0;  // rkf45.py:462-469
        } else {
            L.rejected = true;
        }
        L.h_abs = fabs(h) * factor;
        if (!accept) continue;

        // ---- accepted step ----
        ++L.acc;
        L.t = (t_new != t_new) ? tb : t_new;  // rkf45.py:503
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            L.y[c] = ys[c];
            KS(0, c) = kn[c];  // FSAL: f(y_new) is the next pass's self.f (rkf45.py:378 recomputes the same value)
        }
        L.fresh = true;
        double ysum = L.y[0] + L.y[1] + L.y[2] + L.y[3] + L.y[4];  // np.mean(y) is NaN (rkf45.py:400-402)
        bool dead = (ysum != ysum);
        if (!dead && L.t != tb) continue;

        // ---- output time reached (wr.py:831-885) ----
        int code = 1;  // RWRT_RAY_ALIVE
        if (dead) code = 4;
        if (fabs(L.y[1]) >= HALFPI) {  // wr.py:838-843
            dead = true;
            code = 2;
        }
        if (!dead) {
            double dd = haversine(L.y[0], L.y[1], L.plon, L.plat);  // wr.py:844-850
            if (fabs(dd) >= p.cut) {
                dead = true;
                code = 3;
            }
        }
        if (!dead) {  // a dead ray's remaining rows are NaN already (pre-filled)
            // wr.py:856-865: ug, vg at the output point = what the last stage evaluated at y_new
            if (masked) {
                double2 gv = output_ugvg_slow(p.g, fb, L.y[0], L.y[1], L.y[2], L.y[3]);
                ug = gv.x;
                vg = gv.y;
            }
            write_row(p, L.slot, L.it, L.y[0], L.y[1], L.y[2], L.y[3], L.y[4], ug, vg);
            L.plon = L.y[0];
            L.plat = L.y[1];
            ++L.it;
        }
        if (dead || L.it >= p.nt) {  // ray finished: publish counters, free the lane
            if (p.accepted) p.accepted[L.slot] = L.acc;
            if (p.attempts) p.attempts[L.slot] = L.att;
            if (p.status) p.status[L.slot] = (signed char)code;
            my_total += L.acc;
            L.slot = -1;
        }
    }
#undef KS
    // warp-level reduction of the accepted-step total (the metric), one atomic per warp
    for (int o = 16; o > 0; o >>= 1) my_total += __shfl_down_sync(FULL, my_total, o);
    if (lane_id == 0 && my_total) atomicAdd(p.total_accepted, (unsigned long long)my_total);
}


// ---------------------------------------------------------------------------------------------
// Cooperative-gather integrator (the product path for every grid and ensemble size).
//
// v3 above is bound by the L1TEX wavefront pipe (profiles/r1_v3_*: 92 % of peak): every lane fetches
// its own 4 corners x 96 B with 24 fully divergent LDG.128 per right-hand side.  Here the warp
// fetches together: a quarter-warp copies the four corner points of ONE ray's cell (4 x 96 B) with
// 16-byte cp.async (LDGSTS) -- a handful of wavefronts per instruction instead of ~32 -- straight into a
// padded shared-memory stage (no landing registers, all copies of a stage in flight at once), only for
// the rays that changed cell since their last stage, and one stage AHEAD of use.  Each lane then reads its
// own four points back with conflict-free LDS.128.  Arithmetic and operation order are unchanged.
// Shared memory per warp: 12.5 KB stage + 6.25 KB for K2..K6 -> 76 800 B per block, 3 blocks per SM exactly.
// ---------------------------------------------------------------------------------------------
// The warp gathers the four corner points of a ray's cell straight from the packed POINT table (4 x 96 B; the two
// latitude neighbours are contiguous).  v4-v11 read one 384-byte record of a cell-major COPY of the table instead:
// same number of 16-byte copies per ray and the same speed at 1 degree, but the copy was 4x the size (399 MB at
// 0.25 degree, 25 GB for a 1000-member ensemble) and those configurations fell back to the per-lane kernel.
constexpr int CO_REC_STRIDE = 400;   // 25 x 16 B per staged ray (4 points x 96 B + 16 B holding the fetch-list entry):
                                     // an odd number of 16-byte chunks, so a quarter-warp's LDS.128 / LDGSTS never conflict
constexpr int CO_STAGE_BYTES = 32 * CO_REC_STRIDE;                         // per warp; 4 warps + K2..K6 = 76 800 B = 3 blocks / SM exactly
constexpr int CO_LIST_OFS = 384, CO_LIST_STRIDE = CO_REC_STRIDE;           // fetch-list entry j sits in the pad of row j

// Stage vectors K2..K6 (25 doubles per ray) live in TENSOR MEMORY: lane = thread, 10 x 32-bit columns per
// stage vector, 64 columns per block.  No tensor-core instruction is involved -- tcgen05.ld / .st with the 32x32b
// shape give every thread a private 32-bit x ncol scratch row.  This takes ~45 LDS/STS wavefronts per stage off the
// shared-memory pipe, which is the busiest unit of this kernel (84 % with K2..K6 in shared memory, profiles/r1_v14):
// 56.9 -> 54.1 ms per C2 pass at the same 12 warps per SM.  Using the freed shared memory for a fourth block per SM
// (16 warps at 128 registers) is slower (57.7 ms): RWRT_CO_BLOCKS stays 3.  RWRT_TMEM_KS=0 keeps K2..K6 in shared memory.
#ifndef RWRT_TMEM_KS
#define RWRT_TMEM_KS 1
#endif
constexpr int CO_TMEM_COLS = 64;     // power of two >= 50
#ifndef RWRT_CO_THREADS
#define RWRT_CO_THREADS 128          // threads per block of rk45_integrate_coop (a multiple of 32, <= 128: one TMEM lane per thread)
#endif
constexpr int CO_THREADS = RWRT_CO_THREADS;
#ifndef RWRT_CO_BLOCKS
#define RWRT_CO_BLOCKS 3
#endif
#if RWRT_TMEM_KS
constexpr int CO_BLOCKS_PER_SM = RWRT_CO_BLOCKS;
constexpr int CO_SMEM_BYTES = (CO_THREADS / 32) * CO_STAGE_BYTES + 16;
#else
constexpr int CO_BLOCKS_PER_SM = 3;
constexpr int CO_SMEM_BYTES = (CO_THREADS / 32) * CO_STAGE_BYTES + 5 * 5 * CO_THREADS * (int)sizeof(double);
#endif

// one stage vector (5 doubles = 10 columns) TMEM -> registers; the values may be used after tm_wait_ld(v)
__device__ __forceinline__ void tm_ld5(unsigned taddr, unsigned (&v)[10])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(v[8]), "=r"(v[9]) : "r"(taddr + 8));
}
// the "+r" operands tie the loaded registers to the wait, so that no use of them is scheduled before it
__device__ __forceinline__ void tm_wait_ld(unsigned (&v)[10])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]));
}
// tcgen05.wait::ld waits for ALL earlier tcgen05.ld of the thread, so one wait serves every vector of a combination;
// the others only need the ordering (an empty statement that redefines the registers after the wait).  Each extra
// wait was an issue slot (a NOP carrying the scoreboard wait) in the stage loop.
#ifndef RWRT_TM_ONEWAIT
#define RWRT_TM_ONEWAIT 1
#endif
__device__ __forceinline__ void tm_bind_ld(unsigned (&v)[10])
{
    asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]));
}
__device__ __forceinline__ double tm_val(const unsigned (&v)[10], int c) { return __hiloint2double((int)v[2 * c + 1], (int)v[2 * c]); }
__device__ __forceinline__ void tm_st5(unsigned taddr, const double *k)
{
    unsigned v[10];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        v[2 * c] = (unsigned)__double2loint(k[c]);
        v[2 * c + 1] = (unsigned)__double2hiint(k[c]);
    }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};\n" ::"r"(taddr + 8), "r"(v[8]), "r"(v[9]) : "memory");
}
__device__ __forceinline__ void tm_wait_st()
{
#ifndef RWRT_TM_NOWAITST
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
#endif
}

// sum_{j<S} K_j[c] * coef[j] in j order (K_1 = k1 in registers, K_2.. from tensor memory); S = number of terms
template <int S>
__device__ __forceinline__ void tm_combo(unsigned tk, const double *k1, const double *coef, double *a)
{
    unsigned v[S > 1 ? S - 1 : 1][10];
#pragma unroll
    for (int j = 0; j < S - 1; ++j) tm_ld5(tk + 10 * j, v[j]);
#pragma unroll
    for (int c = 0; c < 5; ++c) a[c] = k1[c] * coef[0];
#pragma unroll
    for (int j = 0; j < S - 1; ++j) {
        if (j == 0 || !RWRT_TM_ONEWAIT) tm_wait_ld(v[j]);
        else tm_bind_ld(v[j]);
#pragma unroll
        for (int c = 0; c < 5; ++c) a[c] = a[c] + tm_val(v[j], c) * coef[j + 1];
    }
}

// The field table must stay in L2 while 23 GB of trajectories stream through it: with the default policy 13 % of
// the gathered sectors missed L2 (v11, cell-major table), and since a warp waits for ~50-120 sectors per stage nearly
// every wait paid a DRAM round trip.  The gathers therefore carry an evict_last policy; the trajectory rows are
// written once and never read back.
#ifndef RWRT_L2_HINT
#define RWRT_L2_HINT 1
#endif
#ifndef RWRT_CP_CG
#define RWRT_CP_CG 0     // 1: the gathers bypass L1 (cp.async.cg: LDGSTS.BYPASS).  Measured: C2 (1 degree) 41.17 -> 41.00 ms, but C5 (0.25
                         // degree) 1258 -> 1343 ms: there the stages of a step straddle cell edges and L1 serves the re-fetches
#endif
__device__ __forceinline__ unsigned long long l2_policy_evict_last()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void cp_async16(unsigned dst_smem, const void *src, unsigned long long pol)
{
#if RWRT_L2_HINT && RWRT_CP_CG
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(dst_smem), "l"(src), "l"(pol));
#elif RWRT_L2_HINT
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(dst_smem), "l"(src), "l"(pol));
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dst_smem), "l"(src));
#endif
}

// Progress flags of the time quanta: release / acquire at GPU scope instead of a full fence on either side (a pick-up
// is a chain of dependent L2 round trips during which the whole warp waits; MEMBAR.GPU was the longest link).
// The reader does NOT use ld.acquire: that compiles to CCTL.IVALL (the whole L1 is invalidated on every pick-up, and at
// 0.25 degree L1 serves a good part of the gathers).  It reads the flags with ld.relaxed.gpu (served by L2) and everything
// the flag guards with ld.global.cg (__ldcg: L2 only, never L1), issued after the branch on the flag value -- no stale L1
// line can be consulted, and a load is not issued before the branch it depends on has resolved.
#ifndef RWRT_TK_OPAQUE
#define RWRT_TK_OPAQUE 1
#endif
#ifndef RWRT_FLAG_ACQUIRE
#define RWRT_FLAG_ACQUIRE 0
#endif
__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
#if RWRT_FLAG_ACQUIRE
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
#else
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
#endif
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory"); }

// The warp-cooperative gather.  A cell is two runs of 192 contiguous bytes in the point table -- points (x0,y0),(x0,y1)
// and, one longitude row further, (x1,y0),(x1,y1) -- staged back to back (c | a | d | b, 4 x 96 B): 24 pieces of 16 B,
// piece p at byte 16 p of the staged record.  A quarter-warp copies one cell with three cp.async per lane:
// instruction i moves pieces 8 i .. 8 i + 7, i.e. 128 contiguous bytes (i = 1: 64 B of each run).  Up to v20 lane q8
// copied half (q8 & 1) of sector i of corner q8 >> 1, four separate 32-byte sectors per instruction;
// tools/experiments/ldgsts_map.cu (the gather alone, 11 cells per warp and stage, 12 warps per SM): 112 -> 73 cycles of
// the L1 / shared-memory pipe per warp and stage, shared-memory wavefronts of the copies 9.4 -> 6.3 per cell, L1 tag
// look-ups 10.5 -> 6 per cell.  Every piece is addressed as point index x 96 B + a compile-time offset from ONE per-lane
// base, so that a copy costs one IMAD.WIDE:
//     piece q8      = base + 96 pi                                  (base = table + 16 q8)
//     piece 8 + q8  = base + 96 (pi + d1) + 32       d1 = 1 for q8 < 4 (byte 128 + 16 q8 of run 0),
//                                                    nlat - 1 for q8 >= 4 (byte 16 (q8 - 4) of run 1)
//     piece 16 + q8 = base + 96 (pi + nlat) + 64     (byte 64 + 16 q8 of run 1)
// Cells with a clamped far index (grid edge, NaN position; interpolation.py:116-119: x1 == x0 makes run 1 a copy of
// run 0, y1 == y0 makes a run's second point its first) never enter the list: their owner lane copies its own 24
// pieces (fetch_self; one in 10^5 stages on a cyclic grid), so the copy loop carries no clamp test (v21 paid it per trip).
// RWRT_PAD_F10 (the 12th double of a table point holds the 11th field of its latitude neighbour, rwrt_bs.cu): the
// owner reads the four 11th fields of its cell with two LDS.128 instead of four LDS.64, which conflict two-way on the
// 400-byte rows (4 wavefronts each, like an LDS.128).  In a y-clamped cell the neighbour IS the point itself, so the
// self copy writes that field into both halves of the piece.
#ifndef RWRT_PAD_F10
#define RWRT_PAD_F10 1
#endif
#ifndef RWRT_FETCH_UNROLL
#define RWRT_FETCH_UNROLL 1          /* 1: not unrolled (38.08 ms on C2); 0: the compiler's choice, four times with remainder code \
                                        (39.83 ms: a quarter-warp has about three list entries); 2: 38.90 ms */
#endif
#ifndef RWRT_COMBO_TREE
#define RWRT_COMBO_TREE 1
#endif
#ifndef RWRT_H_OPAQUE
#define RWRT_H_OPAQUE 1              /* the step h stays in a register pair: 38.09 -> 37.94 ms on C2 */
#endif
#if RWRT_FETCH_UNROLL == 1
#define RWRT_FETCH_LOOP_PRAGMA _Pragma("unroll 1")
#elif RWRT_FETCH_UNROLL == 2
#define RWRT_FETCH_LOOP_PRAGMA _Pragma("unroll 2")
#else
#define RWRT_FETCH_LOOP_PRAGMA
#endif
#define RWRT_FETCH_SETUP                                                                                             \
    const unsigned pf_dst = stage_u32 + (unsigned)(q8 * 16);                                                         \
    const char *pf_src = reinterpret_cast<const char *>(p.g.fpk) + q8 * 16;                                          \
    unsigned pf_d1 = (q8 < 4) ? 1u : (unsigned)p.g.nlat - 1u;                                                        \
    asm volatile("" : "+r"(pf_d1), "+l"(pf_src));                           /* kept whole, not re-derived per copy */
__device__ __forceinline__ void cp_async8(unsigned dst_smem, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst_smem), "l"(src));
}
// the owner lane's own copy of a clamped cell (rare)
__device__ __noinline__ void fetch_self(const double *fpk, int nlat, unsigned key, unsigned dst, unsigned long long l2pol)
{
    const unsigned pi = key & 0x3fffffffu;
    const unsigned xs = (key & 0x40000000u) ? 0u : (unsigned)nlat, ys = (key & 0x80000000u) ? 0u : 1u;
#pragma unroll 1
    for (int pc = 0; pc < 24; ++pc) {
        const int r = pc >= 12 ? 1 : 0, o = (pc - 12 * r) * 16, second = o >= 96 ? 1 : 0;
        const char *src = reinterpret_cast<const char *>(fpk) + (size_t)(pi + (r ? xs : 0u) + (second ? ys : 0u)) * (NPK * 8) +
                          (o - 96 * second);
        if (RWRT_PAD_F10 && o == 80 && ys == 0u) {                         // (f10, f10 of the latitude neighbour = itself)
            cp_async8(dst + 16 * pc, src);
            cp_async8(dst + 16 * pc + 8, src);
        } else {
            cp_async16(dst + 16 * pc, src, l2pol);
        }
    }
}

// Warp-cooperative request for the four corner points of every lane's cell (used by rk45_integrate_coop and
// rk4_integrate_coop; expects the kernels' local names: p, FULL, lane_id, grp, q8, member_pt, prev_key, fetch_list,
// the per-lane copy constants of RWRT_FETCH_SETUP, FETCH_NANSKIP, l2pol and the outputs inr, x0, y0, sx, sy).
#define RWRT_FETCH(LON, LAT)                                                                                         \
    {                                                                                                                \
        inr = fabs(LAT) <= HALFPI;                                          /* bs.py:787 */                          \
        unsigned clamp_;                                                                                             \
        locate_cell<RWRT_FAST, FETCH_NANSKIP>(p.g, LON, inr ? (LAT) : 0.0, x0, y0, clamp_, sx, sy);                                 \
        /* what gets staged = point (x0, y0) AND the clamp pattern (interpolation.py:116-119: x1, y1 come from the   \
           UNCLAMPED index, so a NaN / out-of-range position can pair (x0, y0) with x1 == x0); both are in the key: \
           bits 0..29 point index in the table (all members), bit 30 x clamped, bit 31 y clamped */                  \
        const unsigned key = (member_pt + (unsigned)(x0 * p.g.nlat + y0)) | clamp_;                                  \
        /* a ray often stays in its cell from one stage to the next: its corners are still staged, no new copy */   \
        const bool need = (key != prev_key);                                                                         \
        const bool plain = need && clamp_ == 0u;                                                                     \
        const unsigned need_mask = __ballot_sync(FULL, plain);                                                       \
        prev_key = key;                                                                                              \
        /* compact list of the cells to fetch: (point index, destination lane) */                                    \
        if (plain)                                                                                                   \
            *reinterpret_cast<uint2 *>(fetch_list + __popc(need_mask & ((1u << lane_id) - 1u)) * CO_LIST_STRIDE) =   \
                make_uint2(key, (unsigned)lane_id * CO_REC_STRIDE);                                                  \
        if (need && clamp_ != 0u) fetch_self(p.g.fpk, p.g.nlat, key, stage_u32 + (unsigned)lane_id * CO_REC_STRIDE, l2pol); \
        __syncwarp();                                                                                                \
        const int nfetch = __popc(need_mask);                                                                        \
        /* quarter-warp `grp` copies list entries grp, grp + 4, ... */                                               \
        RWRT_FETCH_LOOP_PRAGMA                                                                                       \
        for (int j = grp; j < nfetch; j += 4) {                                                                      \
            const uint2 e = *reinterpret_cast<const uint2 *>(fetch_list + j * CO_LIST_STRIDE);                       \
            const unsigned dst = pf_dst + e.y;                                                                       \
            cp_async16(dst, pf_src + (size_t)e.x * (NPK * 8), l2pol);                                                \
            cp_async16(dst + 128, pf_src + (size_t)(e.x + pf_d1) * (NPK * 8) + 32, l2pol);                           \
            cp_async16(dst + 256, pf_src + (size_t)(e.x + (unsigned)p.g.nlat) * (NPK * 8) + 64, l2pol);              \
        }                                                                                                            \
        asm volatile("cp.async.commit_group;\n" ::);                                                                 \
    }

// The owner lane's blend of its staged cell (interpolation.py:132-133): a*wa + b*wb + c*wc + d*wd for the 11 fields;
// corners c = (x0,y0) at byte 0, a = (x0,y1) at 96, d = (x1,y0) at 192, b = (x1,y1) at 288; fields 2k, 2k+1 per LDS.128.
#if RWRT_PAD_F10
#define RWRT_BLEND(f, rec, wa, wb, wc, wd)                                                                           \
    _Pragma("unroll") for (int k_ = 0; k_ < 5; ++k_) {                                                               \
        const double2 c2 = *reinterpret_cast<const double2 *>((rec) + k_ * 16);                                      \
        const double2 a2 = *reinterpret_cast<const double2 *>((rec) + 96 + k_ * 16);                                 \
        const double2 d2 = *reinterpret_cast<const double2 *>((rec) + 192 + k_ * 16);                                \
        const double2 b2 = *reinterpret_cast<const double2 *>((rec) + 288 + k_ * 16);                                \
        f[2 * k_] = a2.x * (wa) + b2.x * (wb) + c2.x * (wc) + d2.x * (wd);                                           \
        f[2 * k_ + 1] = a2.y * (wa) + b2.y * (wb) + c2.y * (wc) + d2.y * (wd);                                       \
    }                                                                                                                \
    {   /* 11th field: (c, a) and (d, b) side by side in the sixth piece of c and d */                               \
        const double2 ca = *reinterpret_cast<const double2 *>((rec) + 80);                                           \
        const double2 db = *reinterpret_cast<const double2 *>((rec) + 192 + 80);                                     \
        f[10] = ca.y * (wa) + db.y * (wb) + ca.x * (wc) + db.x * (wd);                                               \
    }
#else
#define RWRT_BLEND(f, rec, wa, wb, wc, wd)                                                                           \
    _Pragma("unroll") for (int k_ = 0; k_ < 6; ++k_) {                                                               \
        const double2 c2 = *reinterpret_cast<const double2 *>((rec) + k_ * 16);                                      \
        const double2 a2 = *reinterpret_cast<const double2 *>((rec) + 96 + k_ * 16);                                 \
        const double2 d2 = *reinterpret_cast<const double2 *>((rec) + 192 + k_ * 16);                                \
        const double2 b2 = *reinterpret_cast<const double2 *>((rec) + 288 + k_ * 16);                                \
        f[2 * k_] = a2.x * (wa) + b2.x * (wb) + c2.x * (wc) + d2.x * (wd);                                           \
        if (2 * k_ + 1 < 11) f[2 * k_ + 1] = a2.y * (wa) + b2.y * (wb) + c2.y * (wc) + d2.y * (wd);                  \
    }
#endif

// OUT: what happens to an output row -- OUT_ROWS (stored: record or planes; the product path of rwrt_ray_run_rk45),
// OUT_GRID (added to the ray-density / flux grids: rwrt_ray_sweep_grid, no trajectory is materialised at all), or both.
// Separate instantiations: the grid arithmetic (two fp64 divisions per output row) stays out of the plain kernel.
#ifdef RWRT_CO_MAXREG   /* occupancy experiments: an explicit register cap instead of the one launch bounds derive */
#define RWRT_CO_BOUNDS __maxnreg__(RWRT_CO_MAXREG)
#else
#define RWRT_CO_BOUNDS __launch_bounds__(CO_THREADS, CO_BLOCKS_PER_SM)
#endif
template <int OUT>
__global__ void RWRT_CO_BOUNDS rk45_integrate_coop(RunParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x;
    int lane_id = tid & 31;
    const int warp_id = tid >> 5;
#if RWRT_TK_OPAQUE
    asm volatile("" : "+r"(lane_id));                                      // opaque: ptxas re-derived it from %tid (S2R, ~25 cycles) in every stage
#endif
    unsigned char *stage = smem_raw + warp_id * CO_STAGE_BYTES;            // this warp's 32 x 384 B
    const unsigned stage_u32 = (unsigned)__cvta_generic_to_shared(stage);
#if RWRT_TMEM_KS
    // tensor-memory scratch for K2..K6: warp 0 allocates 64 columns, every warp addresses its own 32 lanes
    unsigned *tm_slot = reinterpret_cast<unsigned *>(smem_raw + (CO_THREADS / 32) * CO_STAGE_BYTES);
    if (warp_id == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                         (unsigned)__cvta_generic_to_shared(tm_slot)),
                     "n"(CO_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tm_base = *tm_slot;
#if RWRT_TK_OPAQUE
    unsigned tk = tm_base + ((unsigned)(warp_id * 32) << 16);             // + 10 * (s - 1): K_{s+1}, s = 1..5
    asm volatile("" : "+r"(tk));                                           // opaque: kept in a register instead of being recomputed from %tid (S2R) in every stage
#else
    const unsigned tk = tm_base + ((unsigned)(warp_id * 32) << 16);       // + 10 * (s - 1): K_{s+1}, s = 1..5
#endif
#else
    double *ks = reinterpret_cast<double *>(smem_raw + (CO_THREADS / 32) * CO_STAGE_BYTES);
#define KS(s, c) ks[(((s) - 1) * 5 + (c)) * CO_THREADS + tid]            /* K2..K6 <-> s = 1..5 */
#endif
    Lane L;
    L.slot = -1;
    double k1[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 5; ++c) L.y[c] = (c < 2) ? 0.0 : 1.0;              // benign state for idle lanes
    L.h_abs = 1.0;
    L.t = 0.0;
    const double *fb = p.g.fpk;
    bool queue_empty = false;
    long long my_total = 0;
    const int q8 = lane_id & 7, grp = lane_id >> 3;
    double pcos = 1.0;                                                     // cos(latitude of the previous output)
    unsigned prev_key = 0xffffffffu;                                       // what is staged for this lane: point index | clamp pattern
    unsigned member_pt = 0;                                                // first point of this lane's ensemble member
    long long my_ray = -1;                                                 // index of this lane's ray in the state array
    int acc0 = 0;                                                          // accepted steps when this launch picked the ray up
    const double shc = sin(0.5 * p.cut), sin2_half_cut = (p.cut >= PI) ? 2.0 : shc * shc;  // angle <= pi: never cut
    const unsigned char *my_rec = stage + lane_id * CO_REC_STRIDE;        // this lane's staged cell record
    const unsigned long long l2pol = l2_policy_evict_last();
    unsigned char *fetch_list = stage + CO_LIST_OFS;                       // entry j at fetch_list + j * CO_LIST_STRIDE
    RWRT_FETCH_SETUP
    constexpr bool FETCH_NANSKIP = false;                                  // a NaN ray is dropped at its next output row

    // Work items are (time quantum q, ray r), handed out q-major by one global counter: a ray is integrated a quantum
    // of output rows at a time, parked (state saved, progress flag set) and picked up again -- by whichever lane is
    // free -- when its next item comes up.  With whole rays as items the launch ended with a tail of up to one ray
    // length (3.5 ms of serial work on C2; lanes idled 4.5 % of the time); with quanta the tail is one quantum.  No
    // barrier separates the quanta: item (q, r) only waits for item (q - 1, r), which was handed out a whole round
    // earlier.  A waiting lane never spins while other lanes of its warp run (it re-checks once per attempt).
    // The quanta shrink geometrically: quantum q ends where all but 1 / 2^(qshift (q + 1)) of the window's rows are done
    // (with qshift = 2: 3/4 of the rows, then 3/16, 3/64 ...).  Few pick-ups per ray (each is three dependent L2 round trips
    // during which the whole warp waits), yet the LAST items -- the ones that make the tail, and rays differ: 3 % of the
    // 0.25-degree rays take more than twice the mean number of steps, one in 10^6 seven times -- are a few rows long.
    // A lane that holds an item it cannot start yet keeps it in my_ray (L.slot == -2); L.slot == -1: idle without an item.
    const long long nq = p.nquanta > 1 ? p.nquanta : 1;
    const unsigned long long total_items = (unsigned long long)p.nlive * (unsigned long long)nq;
    const int win_len = p.it_end - p.it_begin;
    int it_stop = p.it_end;                                                // this lane's ray parks before output row it_stop
    for (;;) {
        __syncwarp();
        // ---- refill idle lanes from the queue (one atomic per warp) ----
        const bool want = (L.slot == -1);
        const unsigned want_mask = __ballot_sync(FULL, want);
        if (want_mask == FULL && queue_empty) break;                       // no lane has a ray or holds an item, nothing left to draw
        if (want_mask != 0 && !queue_empty) {
            int leader = __ffs(want_mask) - 1;
            unsigned long long base = 0;
            if (lane_id == leader) base = atomicAdd(p.queue, (unsigned long long)__popc(want_mask));
            base = __shfl_sync(FULL, base, leader);
            if (base >= total_items) queue_empty = true;
            if (want) {
                const unsigned long long id = base + __popc(want_mask & ((1u << lane_id) - 1));
                if (id < total_items) {
                    my_ray = (long long)id;
                    L.slot = -2;
                    L.acc = 0;                                             // (idle lanes: counts the attempts spent waiting, see below)
                }
            }
        }
        if (L.slot == -2) {
            const long long q = my_ray / p.nlive, mine = my_ray - q * p.nlive;
            const double *st = p.state + (size_t)mine * RWRT_STATE;
            // both flags in one round trip: finished (sign bit of the attempt counter) / parked at row rows_done
            const int att0 = ld_acquire(reinterpret_cast<const int *>(st + 15) + 1);
            const int rd = ld_acquire(p.rows_done + mine);
            const int b0 = p.it_begin + win_len - (win_len >> (p.qshift * (int)q));            // q = 0: the window start
            const int b1 = (q + 1 >= nq) ? p.it_end : p.it_begin + win_len - (win_len >> (p.qshift * ((int)q + 1)));
            if (att0 < 0 || b1 <= b0) {
                L.slot = -1;                                               // the ray finished / died earlier (or an empty quantum)
            } else if (rd == b0) {   // parked exactly at this quantum's start: its state, written before the flag, is visible (acquire)
                my_ray = mine;
                it_stop = b1;
                L.slot = p.live[mine];
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    L.y[c] = __ldcg(st + c);
                    k1[c] = __ldcg(st + 5 + c);
                }
                L.h_abs = __ldcg(st + 10);
                pcos = __ldcg(st + 11);
                L.t = __ldcg(st + 12);
                L.plon = __ldcg(st + 13);
                L.plat = __ldcg(st + 14);
                L.acc = __ldcg(reinterpret_cast<const int *>(st + 15));
                acc0 = L.acc;
                L.att = __ldcg(reinterpret_cast<const int *>(st + 15) + 1);   // (att0 was read before the flag: only its sign is used)
                L.it = b0;
                L.rejected = false;
                L.fresh = true;
                fb = member_base(p, L.slot);
                member_pt = (unsigned)((fb - p.g.fpk) / NPK);
            } else if (++L.acc > (1 << 21)) {
                // Watchdog: the previous quantum of this ray is at most one quantum of integration away (tens of ms); two
                // million attempts (> 10 s) of waiting can only mean a lost progress flag.  Give the item up and raise the
                // flag the host checks (rwrt_last_run_stats fails) instead of hanging the device.  (No trap here: an exit
                // path in this region costs ptxas its convergence proof for the stage loop, like the early `continue`.)
                p.queue[3] = 1ull;
                L.slot = -1;
            }
        }
        // (A warp whose lanes only WAIT for rays that other warps are still integrating runs an attempt on the benign idle
        // state: an early `continue` / a second exit test here makes ptxas guard every __syncwarp and tcgen05 instruction of
        // the stage loop with BRA.DIV + WARPSYNC, which costs 9 %.)
        const bool active = (L.slot >= 0);

        // ---- one attempt (rkf45.py:410-502); idle lanes run along on a benign state (they serve the gathers) ----
        double tb = L.it * p.tstep;  // wr.py:798-801
        if (L.it == p.nt - 1 && tb > p.ttotal) tb = p.ttotal;
        if (active && L.fresh) {  // rkf45.py:383-387, once per pass
            if (!(L.h_abs >= p.gmin) && L.h_abs == L.h_abs) L.h_abs = p.gmin;
            L.rejected = false;
            L.fresh = false;
        }
        double t_new = L.t + L.h_abs;  // rkf45.py:426-435
        if (t_new - tb > 0) t_new = tb;
#if RWRT_H_OPAQUE
        double h = active ? (t_new - L.t) : 0.0;
        asm volatile("" : "+d"(h));                                        // kept in a register pair: ptxas recomputed it (DADD + 2 ISETP + 2 FSEL) twice per stage
#else
        const double h = active ? (t_new - L.t) : 0.0;
#endif
        double ys[5], kn[5], ug = 0, vg = 0, cos_new = 1.0;
        bool masked = false;
        // Software pipeline over the six stages: the cell record of stage s+1 is requested as soon as stage s
        // has produced its group velocity (d lon/dt, d lat/dt are all the next position needs); the copies fly
        // while stage s finishes dk/dt, dl/dt, dA/dt and stage s+1 evaluates its trigonometry and reciprocals.
        int x0, y0;
        double sx, sy;
        bool inr;
        // stage 1 input (rkf45.py:306-311): y + (K_1 a_21) h
#pragma unroll
        for (int c = 0; c < 5; ++c) ys[c] = L.y[c] + (k1[c] * cC[1][0]) * h;
        __syncwarp();                                                      // the previous attempt's reads of the stage are done
        RWRT_FETCH(ys[0], ys[1])
#pragma unroll 1
        for (int s = 1; s <= 6; ++s) {
            // work that does not need the fields, overlapped with the copies (interpolation.py:112-130 weights)
            const double lat = ys[1];
            const double wa = (1 - sx) * sy, wb = sx * sy, wc = (1 - sx) * (1 - sy), wd = sx * (1 - sy);
            const bool err = (fabs(lat) >= HALFPI) || (fabs(ys[3]) >= 100.0);   // wr.py:508-510
            // the part of the NEXT stage's input that is already known: sum_{j<=s} K_j c_{s+1,j} (j order, rkf45.py:310);
            // this stage's own K_{s+1} c_{s+1,s+1} is added last, below.  Read here, the stage vectors arrive from
            // shared memory under the copy wait instead of on the critical path behind the right-hand side.
            double part[5];
#if RWRT_TMEM_KS
            __syncwarp();
            if (s > 1) tm_wait_st();                                       // K_s of the previous stage has landed
#if RWRT_COMBO_TREE
            // a tree of direct branches instead of the jump table ptxas makes of a switch (LDC + BRX: the warp waits for the
            // constant load, then for the indirect branch to resolve)
            if (s <= 2) {
                if (s < 2) tm_combo<1>(tk, k1, cC[2], part);
                else tm_combo<2>(tk, k1, cC[3], part);
            } else if (s <= 4) {
                if (s < 4) tm_combo<3>(tk, k1, cC[4], part);
                else tm_combo<4>(tk, k1, cC[5], part);
            } else if (s < 6) {
                tm_combo<5>(tk, k1, cC[6], part);
            } else {
#pragma unroll
                for (int c = 0; c < 5; ++c) part[c] = 0.0;
            }
#else
            switch (s) {
                case 1: tm_combo<1>(tk, k1, cC[2], part); break;
                case 2: tm_combo<2>(tk, k1, cC[3], part); break;
                case 3: tm_combo<3>(tk, k1, cC[4], part); break;
                case 4: tm_combo<4>(tk, k1, cC[5], part); break;
                case 5: tm_combo<5>(tk, k1, cC[6], part); break;
                default:
#pragma unroll
                    for (int c = 0; c < 5; ++c) part[c] = 0.0;
                    break;
            }
#endif
#else
#define RWRT_PART(c) \
    switch (s) { \
        case 1: part[c] = k1[c] * cC[2][0]; break; \
        case 2: part[c] = k1[c] * cC[3][0] + KS(1, c) * cC[3][1]; break; \
        case 3: part[c] = k1[c] * cC[4][0] + KS(1, c) * cC[4][1] + KS(2, c) * cC[4][2]; break; \
        case 4: part[c] = k1[c] * cC[5][0] + KS(1, c) * cC[5][1] + KS(2, c) * cC[5][2] + KS(3, c) * cC[5][3]; break; \
        case 5: part[c] = k1[c] * cC[6][0] + KS(1, c) * cC[6][1] + KS(2, c) * cC[6][2] + KS(3, c) * cC[6][3] + KS(4, c) * cC[6][4]; break; \
        default: part[c] = 0.0; break; \
    }
            RWRT_PART(0) RWRT_PART(1) RWRT_PART(2) RWRT_PART(3) RWRT_PART(4)
#undef RWRT_PART
#endif
            const double clast = cC[s < 6 ? s + 1 : 6][s < 6 ? s : 5];      // coefficient of this stage's K_{s+1} in stage s+1's input
            const Trig tg = trig_of_lat<RWRT_FAST>(inr ? lat : 0.0);
            // wr.py:508-510 sets l := NaN under the mask and computes on: every derivative and ug, vg come out NaN
            // (they all depend on l / k).  Same here, by poisoning l -- no branch.
            double yp[5] = {ys[0], ys[1], ys[2], (err || !inr) ? qnan() : ys[3], ys[4]};   // !inr: bs.py:787
            WaveRecip wr;
            if (RWRT_FAST) wr = wave_recip(yp[2], yp[3]);
            // the "+d" operands pin the field-independent results above BEFORE the wait (the compiler is otherwise
            // free to sink that arithmetic below it, which exposes the whole copy latency)
            double wa_ = wa, wd_ = wd, tcp = tg.cp, ttl = tg.tl, wkap = RWRT_FAST ? wr.kap : 0.0, wrden = RWRT_FAST ? wr.rden : 0.0;
            asm volatile("" : "+d"(wa_), "+d"(wd_), "+d"(tcp), "+d"(ttl), "+d"(wkap), "+d"(wrden));
            asm volatile("" : "+d"(part[0]), "+d"(part[1]), "+d"(part[2]), "+d"(part[3]), "+d"(part[4]));
            // ... and the wait sits in a basic block of its own (always taken: nt >= 2 here), or ptxas hoists the
            // DEPBAR to the top of the block to start the record loads early
            if (p.nt > 1) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            __syncwarp();
            Trig tgp = tg;
            tgp.cp = tcp;
            tgp.tl = ttl;
            if (RWRT_FAST) {
                wr.kap = wkap;
                wr.rden = wrden;
            }
            // blend (interpolation.py:132-133): a*wa + b*wb + c*wc + d*wd, fields k = 0..10
            double f[12];
            RWRT_BLEND(f, my_rec, wa_, wb, wc, wd_)
            masked = err;
            Merc m = mercator_from_fields<RWRT_FAST>(f, tgp);
            rhs_position<RWRT_FAST>(m, yp, wr, kn, ug, vg);               // kn[0], kn[1], ug, vg
            cos_new = tgp.cp;
            double yn[5];
            if (s < 6) {
                // next stage's position: y + (sum_{j<=s} K_j c) h with K_{s+1} = kn still in registers
#define RWRT_NEXT(c) yn[c] = L.y[c] + (part[c] + kn[c] * clast) * h;
                RWRT_NEXT(0)
                RWRT_NEXT(1)
                __syncwarp();                                              // every lane has read its record of stage s
                RWRT_FETCH(yn[0], yn[1])
                // keep the rest of this stage's arithmetic BEHIND the copy requests (it all depends on l / k)
                if (RWRT_FAST) asm volatile("" : "+d"(wr.kap), "+d"(wr.rkk));
                else asm volatile("" : "+d"(yp[3]));
            }
            rhs_wavenumber<RWRT_FAST>(m, yp, wr, kn);                     // kn[2], kn[3], kn[4]
            if (s < 6) {
#if RWRT_TMEM_KS
                __syncwarp();
                tm_st5(tk + 10 * (s - 1), kn);
#else
#pragma unroll
                for (int c = 0; c < 5; ++c) KS(s, c) = kn[c];
#endif
                RWRT_NEXT(2)
                RWRT_NEXT(3)
                RWRT_NEXT(4)
#pragma unroll
                for (int c = 0; c < 5; ++c) ys[c] = yn[c];
            }
        }
#undef RWRT_NEXT
        // now ys = y_new and kn = f(y_new) (rkf45.py:317-319); error estimate sum_s K_s E_s (rkf45.py:442-446)
        double ee[5];
#if RWRT_TMEM_KS
        {
            const double cE[6] = {E1, E2, E3, E4, E5, E6};
            __syncwarp();
            tm_wait_st();
            tm_combo<6>(tk, k1, cE, ee);                                   // warp-wide (tcgen05.ld is .sync.aligned)
#pragma unroll
            for (int c = 0; c < 5; ++c) ee[c] = ee[c] + kn[c] * E7;
        }
        if (!active) continue;
#else
        if (!active) continue;
#pragma unroll
        for (int c = 0; c < 5; ++c)
            ee[c] = k1[c] * E1 + KS(1, c) * E2 + KS(2, c) * E3 + KS(3, c) * E4 + KS(4, c) * E5 + KS(5, c) * E6 + kn[c] * E7;
#endif
        ++L.att;
        double r[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) {  // rkf45.py:442-446
            // np.maximum propagates NaN (rkf45.py:443): np_absmax does, in both builds
            double scale = p.atol + np_absmax(L.y[c], ys[c]) * p.rtol;
            const double e = ee[c];
            r[c] = RWRT_FAST ? (h * e) * rcp_fast(scale) : (h * e) / scale;   // scale >= atol > 0, never denormal
        }
        double err = RWRT_FAST ? rms5_fast(r[0], r[1], r[2], r[3], r[4]) : rms5(r[0], r[1], r[2], r[3], r[4]);
        if (err != err) err = 0;
        const bool accept = err < 1;
        double factor = step_factor(err, accept);
        if (accept) {
            if (L.rejected && factor > 1.0) factor = 1.0;  // rkf45.py:462-469
        } else {
            L.rejected = true;
        }
        L.h_abs = fabs(h) * factor;
        if (!accept) continue;

        // ---- accepted step ----
        ++L.acc;
        L.t = (t_new != t_new) ? tb : t_new;  // rkf45.py:503
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            L.y[c] = ys[c];
            k1[c] = kn[c];  // FSAL: f(y_new) is the next pass's self.f (rkf45.py:378 recomputes the same value)
        }
        L.fresh = true;
        double ysum = L.y[0] + L.y[1] + L.y[2] + L.y[3] + L.y[4];  // np.mean(y) is NaN (rkf45.py:400-402)
        bool dead = (ysum != ysum);
        if (!dead && L.t != tb) continue;

        // ---- output time reached (wr.py:831-885) ----
        int code = 1;  // RWRT_RAY_ALIVE
        if (dead) code = 4;
        if (fabs(L.y[1]) >= HALFPI) {  // wr.py:838-843
            dead = true;
            code = 2;
        }
        if (!dead) {
            bool too_fast;
            if (RWRT_FAST && !masked) {
                too_fast = speed_cutoff_fast(L.y[0], L.y[1], L.plon, L.plat, cos_new, pcos, sin2_half_cut);
            } else {
                double dd = RWRT_FAST ? haversine_cold(L.y[0], L.y[1], L.plon, L.plat) : haversine(L.y[0], L.y[1], L.plon, L.plat);  // wr.py:844-850
                too_fast = fabs(dd) >= p.cut;
            }
            if (too_fast) {
                dead = true;
                code = 3;
            }
        }
        if (!dead) {  // a dead ray's remaining rows are NaN already (pre-filled)
            // wr.py:856-865: ug, vg at the output point = what the last stage evaluated at y_new
            if (masked) {
                double2 gv = output_ugvg_slow(p.g, fb, L.y[0], L.y[1], L.y[2], L.y[3]);
                ug = gv.x;
                vg = gv.y;
            }
            if (OUT & OUT_ROWS) {
                if (p.rec) {  // record layout: 64 contiguous bytes per (ray, row), four 16-byte stores
                    double2 *r = reinterpret_cast<double2 *>(p.rec + ((size_t)my_ray * p.rec_rows + (L.it - p.rec_it0)) * 8);
                    r[0] = make_double2(L.y[0], L.y[1]);
                    r[1] = make_double2(L.y[2], L.y[3]);
                    r[2] = make_double2(L.y[4], ug);
                    r[3] = make_double2(vg, 0.0);   // plain stores: streaming (.cs) ones measured the same
                } else {
                    write_row(p, L.slot, L.it, L.y[0], L.y[1], L.y[2], L.y[3], L.y[4], ug, vg);
                }
            }
            if (OUT & OUT_GRID) {  // the sample (lon, lat, l, ug, vg) of this row, exactly as a stored row would hold it
                const long long cell = grid_cell(L.y[0], L.y[1], L.y[3], p.g_dlon, p.g_dlat, p.g_ncx, p.g_ncy);
                if (cell >= 0) grid_add(p.g_count, p.g_fx, p.g_fy, cell, ug, vg);
            }
            L.plon = L.y[0];
            L.plat = L.y[1];
            pcos = masked ? (RWRT_FAST ? cos_cold(L.y[1]) : cos(L.y[1])) : cos_new;
            ++L.it;
        }
        const bool finished = dead || L.it >= p.nt;
        if (finished || L.it >= it_stop) {  // finished, or parked at the end of its time quantum / this time window
            double *st = p.state + (size_t)my_ray * RWRT_STATE;
            if (finished) {
                if (p.accepted) p.accepted[L.slot] = L.acc;
                if (p.attempts) p.attempts[L.slot] = L.att;
                if (p.status) p.status[L.slot] = (signed char)code;
                p.rows_done[my_ray] = L.it;                               // rows 0 .. L.it-1 of this ray are valid
                st_release(reinterpret_cast<int *>(st + 15) + 1, L.att | (int)0x80000000);
            } else {
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    st[c] = L.y[c];
                    st[5 + c] = k1[c];
                }
                st[10] = L.h_abs;
                st[11] = pcos;
                st[12] = L.t;
                st[13] = L.plon;
                st[14] = L.plat;
                reinterpret_cast<int *>(st + 15)[0] = L.acc;
                reinterpret_cast<int *>(st + 15)[1] = L.att;
                st_release(p.rows_done + my_ray, L.it);                   // parked (the state before the flag): the next quantum's item may start
            }
            my_total += L.acc - acc0;
            L.slot = -1;
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                L.y[c] = (c < 2) ? 0.0 : 1.0;
                k1[c] = 0.0;
            }
            L.h_abs = 1.0;
            L.t = 0.0;
            L.it = 1;
        }
    }
#if RWRT_TMEM_KS
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp_id == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm_base), "n"(CO_TMEM_COLS) : "memory");
#else
#undef KS
#endif
    for (int o = 16; o > 0; o >>= 1) my_total += __shfl_down_sync(FULL, my_total, o);
    if (lane_id == 0 && my_total) atomicAdd(p.total_accepted, (unsigned long long)my_total);
}

// ---------------------------------------------------------------------------------------------
// Fixed-step RK4 path: WR.core_ray_run_numpy (wr.py:702-765) with rk4_step_numpy (wr.py:583-622) and
// core_rk4_step (wr.py:89-95) -- the reference's DEFAULT integrator (inte_method='').  One slot per
// thread, every thread takes the same nt-1 steps, so rows are written fully coalesced; all slots run
// (NaN-root slots are not masked by the reference and turn all-NaN at row 1).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) rk4_integrate(RunParams p)
{
    long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= p.N) return;
    const size_t pl = (size_t)p.nt * p.N;
    const double *fb = member_base(p, slot);
    double y[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) y[c] = p.out[c * pl + slot];
    double plon = y[0], plat = y[1];
    const double dt = p.tstep;
    int code = (y[3] == y[3]) ? 1 : 0;
    for (int it = 0; it < p.nt - 1; ++it) {
        double k1[5], k2[5], k3[5], k4[5], ys[5], yn[5], ug, vg;
        bool m1, m2, m3, m4;
        ray_rhs<RWRT_FAST>(p.g, fb, y, k1, ug, vg, m1);
#pragma unroll
        for (int c = 0; c < 5; ++c) yn[c] = y[c];
        if (!m1) {  // wr.py:598-618
#pragma unroll
            for (int c = 0; c < 5; ++c) ys[c] = y[c] + 0.5 * dt * k1[c];
            ray_rhs<RWRT_FAST>(p.g, fb, ys, k2, ug, vg, m2);
#pragma unroll
            for (int c = 0; c < 5; ++c) ys[c] = y[c] + 0.5 * dt * k2[c];
            ray_rhs<RWRT_FAST>(p.g, fb, ys, k3, ug, vg, m3);
#pragma unroll
            for (int c = 0; c < 5; ++c) ys[c] = y[c] + dt * k3[c];
            ray_rhs<RWRT_FAST>(p.g, fb, ys, k4, ug, vg, m4);
            if (!m2 && !m3 && !m4) {
#pragma unroll
                for (int c = 0; c < 5; ++c) yn[c] = y[c] + (dt / 6.0) * (k1[c] + 2 * k2[c] + 2 * k3[c] + k4[c]);
            }
        }
        if (fabs(yn[1]) >= HALFPI) {  // wr.py:721-726
#pragma unroll
            for (int c = 0; c < 5; ++c) yn[c] = qnan();
            if (code == 1) code = 2;
        }
        double dd = haversine(yn[0], yn[1], plon, plat);  // wr.py:728-734
        if (fabs(dd) >= p.cut) {
#pragma unroll
            for (int c = 0; c < 5; ++c) yn[c] = qnan();
            if (code == 1) code = 3;
        }
        if (code == 1 && !(yn[0] + yn[1] + yn[2] + yn[3] + yn[4] == yn[0] + yn[1] + yn[2] + yn[3] + yn[4])) code = 4;
        Merc m = sample_mercator<RWRT_FAST>(p.g, fb, yn[0], yn[1]);  // wr.py:737-746
        ugvg_extent(m.u, m.v, m.qx, m.qy, yn[2], yn[3], ug, vg);
        write_row(p, slot, it + 1, yn[0], yn[1], yn[2], yn[3], yn[4], ug, vg);
#pragma unroll
        for (int c = 0; c < 5; ++c) y[c] = yn[c];
        plon = yn[0];
        plat = yn[1];
    }
    if (p.status) p.status[slot] = (signed char)code;
}

// ---------------------------------------------------------------------------------------------
// Fixed-step RK4 on the cooperative-gather machinery (the product path of rwrt_ray_run_rk4; rk4_integrate above is
// kept as the cross-check).  Same arithmetic as rk4_integrate, two structural differences:
//   * FSAL: the evaluation at y_new that the output needs (ug, vg; wr.py:737-746) is also the next step's k1
//     (wr.py:598 recomputes the same value) -- 4 right-hand sides per step instead of 5;
//   * the four corner points of each evaluation arrive through the warp's shared staging rows, requested one
//     evaluation ahead (as in rk45_integrate_coop).
// One slot per thread, all slots in lock step: rows are written fully coalesced in the reference layout.
// ---------------------------------------------------------------------------------------------
constexpr int RK4_SMEM_BYTES = (RK_THREADS / 32) * CO_STAGE_BYTES;
__global__ void __launch_bounds__(RK_THREADS, 3) rk4_integrate_coop(RunParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x;
    const int lane_id = tid & 31, warp_id = tid >> 5;
    unsigned char *stage = smem_raw + warp_id * CO_STAGE_BYTES;
    const unsigned stage_u32 = (unsigned)__cvta_generic_to_shared(stage);
    const int q8 = lane_id & 7, grp = lane_id >> 3;
    const unsigned char *my_rec = stage + lane_id * CO_REC_STRIDE;
    const unsigned long long l2pol = l2_policy_evict_last();
    unsigned char *fetch_list = stage + CO_LIST_OFS;
    RWRT_FETCH_SETUP
    constexpr bool FETCH_NANSKIP = true;                                   // no-root slots run along as NaN for all nt - 1 steps
    unsigned prev_key = 0xffffffffu;

    long long slot = (long long)blockIdx.x * blockDim.x + tid;
    const bool mine = slot < p.N;
    if (!mine) slot = p.N - 1;                     // the warp works together: surplus lanes shadow the last slot, write nothing
    const size_t pl = (size_t)p.nt * p.N;
    const double *fb = member_base(p, slot);
    const unsigned member_pt = (unsigned)((fb - p.g.fpk) / NPK);
    double y[5], acc[5], kn[5], ys[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) ys[c] = y[c] = p.out[c * pl + slot];
    double plon = y[0], plat = y[1];
    const double dt = p.tstep;
    int code = (y[3] == y[3]) ? 1 : 0;
    int x0, y0;
    double sx, sy, ug = 0, vg = 0;
    bool inr, m_first = false, m_any = false;
    // evaluation e of step it: e = 0 -> f(y) (only before the first step; afterwards FSAL), 1..3 -> k2..k4, 4 -> f(y_new)
    RWRT_FETCH(ys[0], ys[1])
    int e = 0, it = 0;
    for (;;) {
        const double lat = ys[1];
        const double wa = (1 - sx) * sy, wb = sx * sy, wc = (1 - sx) * (1 - sy), wd = sx * (1 - sy);
        const bool err = (fabs(lat) >= HALFPI) || (fabs(ys[3]) >= 100.0);   // wr.py:508-510
        const Trig tg = trig_of_lat<RWRT_FAST>(inr ? lat : 0.0);
        double yp[5] = {ys[0], ys[1], ys[2], (err || !inr) ? qnan() : ys[3], ys[4]};
        WaveRecip wr;
        if (RWRT_FAST) wr = wave_recip(yp[2], yp[3]);
        double wa_ = wa, wd_ = wd, tcp = tg.cp, ttl = tg.tl, wkap = RWRT_FAST ? wr.kap : 0.0, wrden = RWRT_FAST ? wr.rden : 0.0;
        asm volatile("" : "+d"(wa_), "+d"(wd_), "+d"(tcp), "+d"(ttl), "+d"(wkap), "+d"(wrden));
        if (p.nt > 1) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncwarp();
        Trig tgp = tg;
        tgp.cp = tcp;
        tgp.tl = ttl;
        if (RWRT_FAST) {
            wr.kap = wkap;
            wr.rden = wrden;
        }
        double f[12];
        RWRT_BLEND(f, my_rec, wa_, wb, wc, wd_)
        Merc m = mercator_from_fields<RWRT_FAST>(f, tgp);
        rhs_position<RWRT_FAST>(m, yp, wr, kn, ug, vg);
        rhs_wavenumber<RWRT_FAST>(m, yp, wr, kn);
        // ---- what this evaluation was, and where the next one is ----
        bool done = false;
        if (e == 0 || e == 4) {                     // kn = f(y): k1 of the coming step (wr.py:598)
            if (e == 4) {                           // ... and the output of the step that just ended (wr.py:737-746)
                if (err) {
                    double2 gv = output_ugvg_slow(p.g, fb, y[0], y[1], y[2], y[3]);
                    ug = gv.x;
                    vg = gv.y;
                }
                if (mine) write_row(p, slot, it + 1, y[0], y[1], y[2], y[3], y[4], ug, vg);
                plon = y[0];
                plat = y[1];
                ++it;
                done = it >= p.nt - 1;
            }
            m_first = err;
            m_any = false;
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                acc[c] = kn[c];
                ys[c] = y[c] + 0.5 * dt * kn[c];
            }
            e = 1;
        } else if (e == 1 || e == 2) {              // k2, k3
            m_any |= err;
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                acc[c] = acc[c] + 2 * kn[c];
                ys[c] = (e == 1) ? y[c] + 0.5 * dt * kn[c] : y[c] + dt * kn[c];
            }
            ++e;
        } else {                                    // k4: finish the step (wr.py:598-618, 721-734)
            m_any |= err;
            if (!m_first && !m_any) {
#pragma unroll
                for (int c = 0; c < 5; ++c) y[c] = y[c] + (dt / 6.0) * (acc[c] + kn[c]);
            }
            if (fabs(y[1]) >= HALFPI) {
#pragma unroll
                for (int c = 0; c < 5; ++c) y[c] = qnan();
                if (code == 1) code = 2;
            }
            double dd = haversine(y[0], y[1], plon, plat);
            if (fabs(dd) >= p.cut) {
#pragma unroll
                for (int c = 0; c < 5; ++c) y[c] = qnan();
                if (code == 1) code = 3;
            }
            if (code == 1 && !(y[0] + y[1] + y[2] + y[3] + y[4] == y[0] + y[1] + y[2] + y[3] + y[4])) code = 4;
#pragma unroll
            for (int c = 0; c < 5; ++c) ys[c] = y[c];
            e = 4;
        }
        if (done) break;
        __syncwarp();                               // every lane has read its staged points
        RWRT_FETCH(ys[0], ys[1])
    }
    if (mine && p.status) p.status[slot] = (signed char)code;
}

#undef RWRT_FETCH
#undef RWRT_FETCH_SETUP
#undef RWRT_BLEND

}  // namespace RWRT_VARIANT
}  // namespace rwrt
