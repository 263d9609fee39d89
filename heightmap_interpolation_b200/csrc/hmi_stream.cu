// Streaming temporal-blocked sweep kernel for the symmetric stencils (harmonic, CCST/biharmonic)
// under the reflect-101 border rule — the hot kernel of the path.
//
// Design (B200-first, no counterpart in the reference, which runs one full-grid NumPy/OpenCV pass
// per operator: fd_pde_inpainter.py:311-326):
//   * One WARP owns a column strip of 32*V cells and streams down a chunk of rows.  Each lane holds
//     V consecutive cells of the current row in registers; x-neighbours come from warp shuffles,
//     y-neighbours from the previous rows kept in registers.  No block barrier on the hot loop.
//   * Rows are staged through a per-warp shared-memory ring filled by asynchronous 16-byte copies
//     (cp.async / LDGSTS; with V*sizeof(T) == 16 every copy instruction moves one fully coalesced
//     512-byte row segment) issued D-1 rows ahead, so several KiB per warp are in flight without
//     holding registers or coupling to the load scoreboard.
//   * Two code paths per instantiation: interior strips run a lean loop; strips that touch a grid
//     edge (mirrored ghost columns / rows) run the general one.
//   * Temporal blocking: K relaxation sweeps are fused as a K-stage software pipeline in
//     registers.  Row y enters stage 0 as it is loaded; stage s emits sweep s+1 of row y-r(s+1)
//     (r = stencil radius: 1 harmonic, 2 CCST).  Only the K-th iterate is written back, so HBM
//     traffic per cell update falls from 17/9 B (fp64/fp32) to about 17/K / 9/K B.
//   * Strips overlap by 2*r*K columns and chunks by 2*r*K rows; the overlap cells are recomputed
//     redundantly and never stored (overlapped tiling), so warps are fully independent.
//   * Reflect-101 borders: out-of-grid ghost cells are loaded from their mirror cell (value and
//     mask).  Both stencils are symmetric, so a ghost cell evolves exactly like its mirror image and
//     the mirrored load is equivalent to re-reflecting every intermediate field.
//   * The mask travels down the pipeline as a bit history (V bits per row per lane).
//   * The convergence sums of the last fused sweep are accumulated per lane in fp64 and folded by
//     warp shuffles + one block reduction + a fixed-order final fold (no extra pass over the grid).
#include <cstdio>
#include <cstdlib>

#include "hmi_common.cuh"
#include "hmi_stream.cuh"
#include "hmi_stream_common.cuh"

// build-time tuning knobs (defaults measured on B200, see profiles/)
#ifndef HMI_STREAM_RING
#define HMI_STREAM_RING 8   // ring slots per warp
#endif
#ifndef HMI_STREAM_WPB
#define HMI_STREAM_WPB 4    // warps per CTA
#endif
#ifndef HMI_STREAM_MINB
#define HMI_STREAM_MINB 1   // __launch_bounds__ min CTAs per SM (1 = let ptxas choose)
#endif
#ifndef HMI_STREAM_PEEL
#define HMI_STREAM_PEEL 1   // 1 = interior strips run a steady-state row loop without fetch / store range tests
#endif
#ifndef HMI_STREAM_SKEW
#define HMI_STREAM_SKEW 1   // 1 = stage s+1 consumes stage s's row one step later (independent stages)
#endif

namespace hmi {

// Geometry of one instantiation
template <typename T, int V, int K, int METHOD>
struct Geo {
    static constexpr int N = Chunk16<T>::N;             // cells per 16-byte chunk
    static constexpr int NCH = V / N;                   // 16-byte chunks per lane per row
    static constexpr int R = (METHOD == HMI_CCST) ? 2 : 1;
    static constexpr int H = K * R;                     // rows/columns of halo one launch consumes
    // Skewed pipeline: the row a stage emits is consumed by the next stage in the *next* row step,
    // so the K stages of one step have no data dependence on each other and their fp64 chains
    // interleave (the unskewed loop could issue only every ~4th cycle per warp: each stage waited
    // on the previous one).  Costs K-1 extra rows of latency per chunk, no registers.
    // Measured (profiles/r02_tune_skew_ab.log): +9..13 % for CCST fp64, whose stage is two chained
    // 5-point passes; -5..8 % for harmonic and for fp32, whose stages are too short to need it.
    static constexpr int SK = (HMI_STREAM_SKEW && K > 1 && METHOD == HMI_CCST && sizeof(T) == 8) ? 1 : 0;
    static constexpr int LAG = H + SK * (K - 1);        // step y emits sweep K of row y - LAG
    static constexpr int HV = ((H + V - 1) / V) * V;    // column halo rounded to whole lanes
    static constexpr int S = 32 * V - 2 * HV;           // columns a warp produces
    static constexpr int ROW_BYTES = 32 * V * (int)sizeof(T);
    static constexpr int SLOT = ROW_BYTES + 128;        // one row segment + one mask word per lane
    // TMA staging: the row segment and the 16-byte-aligned superset of its mask bytes land as two
    // bulk copies; D mbarriers per warp follow the ring
    static constexpr int MB_T = ((32 * V + 16 + 15) / 16) * 16;
    static constexpr int SLOT_T = ROW_BYTES + MB_T;
    // 2-D tensor-map staging (interior strips): boxes of RB rows x 32*V columns of f and of the mask,
    // NBX boxes in flight per warp, one mbarrier per box.  RB = 3 matches the x3 unrolled row loop, so
    // the wait sits in phase 0 and the refill in phase 2 with no run-time row counter.
    static constexpr int RB = 3, NBX = 4;
    static constexpr int FBOX = RB * ROW_BYTES;
    // a box must start on a 16-byte boundary in global memory (measured: an odd start faults with "illegal
    // instruction", tools/probes/tma2d_probe.cu); strips start at even columns, which is enough for the
    // cells but not for the 1-byte mask, so the mask box starts at the 16-aligned column below x0 and is
    // 16 bytes wider
    static constexpr int MROW = 32 * V + 16;
    static constexpr int MBOX = ((RB * MROW + 127) / 128) * 128;
    static constexpr int RING_T2 = ((NBX * (FBOX + MBOX) + NBX * 8 + 127) / 128) * 128;
    static_assert(V % N == 0 && NCH >= 1, "a lane moves whole 16-byte chunks");
    static_assert(V == 2 || V == 4, "a lane's mask bytes must sit inside one aligned 4-byte word");
    static_assert(S > 0, "temporal block too deep for this strip width");
};

template <bool SMALL> struct MaskHist { typedef unsigned long long type; };
template <> struct MaskHist<true> { typedef unsigned type; };

// One warp streams its strip down one chunk of rows.  MODE 2 = the strip touches a grid edge (ghost
// columns to mirror, rows to reflect or clamp); interior strips skip all of that and address
// memory incrementally.  NORM = accumulate the convergence sums of the (single) sweep.
// BRD = border rule: HMI_BORDER_REFLECT101, or HMI_BORDER_ZERO (the SciPy convolvers, convolver.py:17-18, 47-54):
// every filtered array — f and, CCST, h = L f — reads zeros outside the grid.  Ghost cells of f then are
// "known cells of value 0" (they stay 0 through all fused sweeps) and ghost cells of h are forced to 0.
template <typename T, int V, int K, int METHOD, int D, int MODE, bool NORM, int STG, int BRD>
__device__ __forceinline__ void stream_rows(const StreamArgs<T> &a, const unsigned ring_s, const int lane,
                                            const int win, const int y0, const int y1, double &n_d2, double &n_f2,
                                            double &n_mx, double &n_nan)
{
    typedef Geo<T, V, K, METHOD> G;
    typedef typename Chunk16<T>::type C16;
    // row staging: 0 = per-lane cp.async ring, 1 = one 1-D bulk copy per row, 2 = 2-D tensor-map boxes
    constexpr bool TMA = (STG == 1), T2 = (STG == 2);
    constexpr int N = G::N, NCH = G::NCH, R = G::R, H = G::H, HV = G::HV, S = G::S;
    constexpr int ROW_BYTES = G::ROW_BYTES, SLOT = TMA ? G::SLOT_T : G::SLOT;
    constexpr bool CC = (METHOD == HMI_CCST);
    constexpr unsigned VMASK = (1u << V) - 1u;
    constexpr int SK = G::SK, LAG = G::LAG;
    constexpr int RB = G::RB, NBX = G::NBX, FBOX = G::FBOX, MBOX = G::MBOX, MROW = G::MROW;
    static_assert(!T2 || MODE == 0, "tensor-map boxes serve interior strips only");
    constexpr bool ZERO = (BRD == HMI_BORDER_ZERO) && MODE >= 1;
    static_assert(BRD == HMI_BORDER_REFLECT101 || (BRD == HMI_BORDER_ZERO && STG == 0), "zero padding uses the cp.async staging");
    static_assert((LAG + 1) * V <= 64, "mask history does not fit 64 bits");
    static_assert(!NORM || K == 1, "the convergence sums are fused into single-sweep launches only");

    // MODE 0: interior (lean loop, incremental addressing); 1: only the rows need care (reflect-101 at
    // the global top/bottom, clamping to the slab's memory) — same compute path as 0; 2: ghost
    // columns as well (mirrored out of other lanes' ring slots).
    constexpr bool REDGE = MODE >= 1;
    constexpr bool CEDGE = MODE == 2 || (TMA && MODE >= 1);
    const int x0 = win * S - HV;
    const int col0 = x0 + lane * V;
    const int ybeg = y0 - H, yend = y1 + H;        // rows streamed: [ybeg, yend)
    const int ystop = yend + SK * (K - 1);         // row steps: [ybeg, ystop) (the skew drains K-1 steps)
    const bool out_lane = (lane >= HV / V) && (lane < 32 - HV / V) && (col0 < a.cols);
    const T dt = a.dt;
    // CCST with folded coefficients: f' = f + c2*(sum of h's four neighbours) + c3*h,
    //   c2 = -dt(1-t), c3 = dt*t - 4*c2      (== f + dt*(t*h - (1-t)*L h))
    const T c2 = -(dt * a.one_minus_tension);
    const T c3 = fma(T(-4), c2, dt * a.tension);

    // Each lane copies the aligned 4-byte mask word that holds its V mask bytes (V = 2: two lanes
    // fetch the same word), so no lane ever reads ring bytes another lane wrote — except the edge
    // path, which mirrors ghost columns out of other lanes' slots.
    const int mcol = col0 & ~3;
    const unsigned msh = (unsigned)(col0 & 3) * 8u;
    const bool f_copies = CEDGE ? ((col0 >= 0) && (col0 + V <= (int)a.pitch)) : true;
    const bool m_copies = CEDGE ? ((mcol >= 0) && (mcol + 4 <= (int)a.mpitch)) : true;

    // Pipeline state.  Stage s keeps a 3-row window of its input field in F[s][0..2]; the row loop
    // is unrolled by 3 so the window rotates by renaming (no register moves): in phase PH the rows
    // above / centre / new sit in slots PH, PH+1, PH+2 (mod 3).  Stage s writes its output row
    // straight into the new-row slot of stage s+1.
    T F[K][3][V];
    T Hh[CC ? K : 1][3][V];                        // CCST: window of h = L f
    T FL[K], FR[K], HL[CC ? K : 1], HR[CC ? K : 1];
    // mask history: V bits per row, newest row in the low bits; 32 bits when they suffice (the bit
    // tests then compile to one LOP3 each instead of a 64-bit shift + compare)
    typedef typename MaskHist<((LAG + 1) * V <= 32) && !TMA && V == 2>::type MH;   // (ptxas runs out of predicate registers on some V = 4 and TMA variants)
    MH mhist = 0;
#pragma unroll
    for (int s = 0; s < K; ++s) {
        FL[s] = FR[s] = T(0);
        if (CC) HL[s] = HR[s] = T(0);
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int v = 0; v < V; ++v) {
                F[s][q][v] = T(0);
                if (CC) Hh[s][q][v] = T(0);
            }
    }

    const unsigned ring_f = ring_s + lane * 16;                // this lane's first 16-byte chunk
    const unsigned ring_m = ring_s + ROW_BYTES + lane * 4;     // this lane's mask word
    const T *gbase = a.src + col0;
    const uint8_t *mbase = a.mask + mcol;
    // running state of the copy engine: next row to fetch and where
    int ynext = ybeg;
    unsigned wofs = 0;                                         // ring byte offset of the slot to fill
    const T *gnext = gbase + (long long)ybeg * a.pitch;        // interior path: incremental
    const uint8_t *mnext = mbase + (long long)ybeg * a.mpitch;

    // TMA variant: lane 0 moves the whole row segment and its mask bytes with two bulk copies that
    // complete on the slot's mbarrier
    const unsigned bar_s = ring_s + D * SLOT;                  // D mbarriers after the ring
    const int mx16 = x0 & ~15;                                 // mask copy starts 16-byte aligned
    unsigned widx = 0;                                         // slot index being filled
    if (TMA) {
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < D; ++i) mbar_init(bar_s + 8 * i, 1);
            mbar_fence_init();
        }
        __syncwarp();
    }

    // T2: lane 0 fetches box `bi` (rows ybeg + RB*bi ...) of f and of the mask into slot `slot`
    const unsigned t2_f = ring_s, t2_m = ring_s + NBX * FBOX, t2_bar = ring_s + NBX * (FBOX + MBOX);
    auto issue_box = [&](const int bi, const unsigned slot) {
        const int ybox = ybeg + RB * bi;
        if (ybox < yend && lane == 0) {
            const unsigned bar = t2_bar + 8 * slot;
            fence_proxy_async();                   // the slot's previous contents were read with ld.shared
            mbar_expect_tx(bar, (unsigned)(FBOX + RB * MROW));
            tma_load_2d(t2_f + slot * FBOX, &a.tmap_f, x0, ybox + a.phys_row0, bar);
            tma_load_2d(t2_m + slot * MBOX, &a.tmap_m, x0 & ~15, ybox + a.phys_row0, bar);
        }
    };
    if (T2) {
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < NBX; ++i) mbar_init(t2_bar + 8 * i, 1);
            mbar_fence_init();
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NBX; ++i) issue_box(i, (unsigned)i);
    }
    int t2_box = 0;                                            // box being consumed
    unsigned t2_slot = 0, t2_par = 0;                          // ... its slot and mbarrier parity

    auto issue = [&](auto steady, auto in_loop) {
        constexpr bool STEADY = decltype(steady)::value != 0;   // the caller guarantees ynext < yend
        constexpr bool IN_LOOP = decltype(in_loop)::value != 0; // the caller sets wofs (the slot read one step ago)
        if (T2) return;
        if (TMA) {
            if (ynext < yend && lane == 0) {
                int ry = ynext;
                int c_lo = x0, c_hi = x0 + 32 * V, m_lo = mx16, m_hi = ((x0 + 32 * V + 15) & ~15);
                if (REDGE) {
                    if (ry < 0 && a.top_border) ry = -ry;
                    else if (ry >= a.rows && a.bottom_border) ry = 2 * (a.rows - 1) - ry;
                    ry = max(a.read_lo, min(a.read_hi - 1, ry));
                    c_lo = max(c_lo, 0); c_hi = min(c_hi, (int)a.pitch);
                    m_lo = max(m_lo, 0); m_hi = min(m_hi, (int)a.mpitch);
                }
                const unsigned bar = bar_s + 8 * widx;
                const unsigned fb = (unsigned)(c_hi - c_lo) * (unsigned)sizeof(T), mbytes = (unsigned)(m_hi - m_lo);
                fence_proxy_async();               // the slot's previous contents were read with ld.shared
                mbar_expect_tx(bar, fb + mbytes);
                tma_load_1d(ring_s + wofs + (unsigned)(c_lo - x0) * (unsigned)sizeof(T),
                            a.src + (long long)ry * a.pitch + c_lo, fb, bar);
                tma_load_1d(ring_s + wofs + ROW_BYTES + (unsigned)(m_lo - mx16),
                            a.mask + (long long)ry * a.mpitch + m_lo, mbytes, bar);
            }
            widx = (widx + 1 == D) ? 0u : widx + 1;
        } else {
            if (STEADY || ynext < yend) {
                const T *prow = gnext;
                const uint8_t *mrow = mnext;
                if (REDGE) {                       // reflect-101 at global edges, clamp to memory
                    int ry = ynext;
                    if (ZERO) ;                        // rows outside the grid are zeroed when they are consumed
                    else if (ry < 0 && a.top_border) ry = -ry;
                    else if (ry >= a.rows && a.bottom_border) ry = 2 * (a.rows - 1) - ry;
                    ry = max(a.read_lo, min(a.read_hi - 1, ry));
                    prow = gbase + (long long)ry * a.pitch;
                    mrow = mbase + (long long)ry * a.mpitch;
                }
                if (f_copies) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) cp_async16(ring_f + wofs + c * 512, prow + c * N);
                }
                if (m_copies) cp_async4(ring_m + wofs, mrow);
            }
            cp_async_commit();
            if (!REDGE) { gnext += a.pitch; mnext += a.mpitch; }
        }
        ++ynext;
        if (TMA || !IN_LOOP) wofs = (wofs + SLOT == D * SLOT) ? 0u : wofs + SLOT;
    };

    if (!T2) {
#pragma unroll
        for (int u = 0; u < D - 1; ++u) issue(Phase<0>(), Phase<0>());
    }

    unsigned rofs = 0;                                         // ring byte offset of the slot to read
    unsigned pofs = (D - 1) * SLOT;                            // ... of the slot read one step ago (refilled next)
    unsigned ridx = 0, rphase = 0;                             // TMA: slot index and mbarrier parity
    T *dptr = a.dst + (long long)(ybeg - LAG) * a.pitch + col0;   // row emitted by step ybeg

    // STEADY steps (interior strips, cp.async staging): the row to fetch exists and the row emitted lies in
    // [y0, y1), so neither is tested
    auto step = [&](auto ph, auto steady, const int y) {
        constexpr int PH = decltype(ph)::value;
        constexpr bool STEADY = decltype(steady)::value != 0;
        constexpr int IU = PH, IC = (PH + 1) % 3, IN = (PH + 2) % 3;
        if (T2) {
            if (PH == 0 && y < yend) mbar_wait(t2_bar + 8 * t2_slot, t2_par);   // the box rows y, y+1, y+2 sit in has landed
        } else if (TMA) {
            if (y < yend) mbar_wait(bar_s + 8 * ridx, rphase);   // both bulk copies of row y have landed
        } else {                                   // (no row is fetched for the drain steps of the skew)
            cp_async_wait<D - 2>();                // this lane's copies of row y have landed
            if (CEDGE) __syncwarp();               // ... and the other lanes' (for the mirrors)
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            T e[N];
            lds16(T2 ? t2_f + t2_slot * FBOX + PH * ROW_BYTES + lane * (16 * NCH) + c * 16
                     : (TMA ? ring_s + rofs + lane * (16 * NCH) + c * 16 : ring_f + rofs + c * 512), e);
#pragma unroll
            for (int q = 0; q < N; ++q) F[0][IN][c * N + q] = e[q];
        }
        // mask bytes are 0/1: one multiply packs them into V bits, cell v -> bit V-1-v
        unsigned mb;
        if (T2) {
            // the mask box starts at column x0 & ~15, so this lane's V mask bytes sit at byte (x0 & 15) + V*lane of the row
            const unsigned ma = t2_m + t2_slot * MBOX + PH * MROW + (unsigned)(x0 & 15) + lane * V;
            if (V == 4) mb = (lds32(ma) * 0x08040201u) >> 24;
            else mb = ((lds16u(ma) * 0x0201u) >> 8) & 3u;
            if (PH == 2) {                         // last row of the box: refill its slot with box t2_box + NBX
                __syncwarp();                      // every lane is done with the slot
                issue_box(t2_box + NBX, t2_slot);
                t2_box += 1;
                t2_slot = (t2_slot + 1 == NBX) ? 0u : t2_slot + 1;
                if (t2_slot == 0) t2_par ^= 1u;
            }
        } else {
            unsigned w, sh = msh;
            if (TMA) {
                const unsigned off = (unsigned)(x0 - mx16) + (unsigned)(V * lane);
                w = lds32(ring_s + rofs + ROW_BYTES + (off & ~3u));
                sh = (off & 3u) * 8u;
            } else {
                w = lds32(ring_m + rofs);
            }
            if (V == 4) mb = (w * 0x08040201u) >> 24;
            else mb = ((((w >> sh) & 0xffffu) * 0x0201u) >> 8) & 3u;
        }
        if (ZERO) {
            // zero padding: cells outside the grid hold 0 and count as known, so every fused sweep keeps them 0
            const bool ghost_row = (y < 0 && a.top_border) || (y >= a.rows && a.bottom_border);
            if (CEDGE) mb &= VMASK;
#pragma unroll
            for (int q = 0; q < V; ++q) {
                const int c = col0 + q;
                if (ghost_row || (CEDGE && (c < 0 || c >= a.cols))) {
                    F[0][IN][q] = T(0);
                    mb |= 1u << (V - 1 - q);
                }
            }
        } else if (CEDGE) {
            // Ghost columns take the value and mask of their mirror cell, which sits in this
            // strip's ring slot (copied by another lane).  Lanes outside the allocation copied
            // nothing and read stale bytes, so clamp the bit group first.
            mb &= VMASK;
#pragma unroll
            for (int q = 0; q < V; ++q) {
                const int c = col0 + q;
                if (c < 0 || c >= a.cols) {
                    const int cm = (c < 0) ? -c : 2 * (a.cols - 1) - c;
                    const int pos = max(0, min(32 * V - 1, cm - x0));
                    const int pl = pos / V, pq = pos % V;
                    unsigned kb;
                    if (TMA) {
                        F[0][IN][q] = lds1(ring_s + rofs + (unsigned)pos * (unsigned)sizeof(T), T(0));
                        kb = lds8(ring_s + rofs + ROW_BYTES + (unsigned)(x0 - mx16) + (unsigned)pos) & 1u;
                    } else {
                        T e[N];
                        lds16(ring_s + rofs + (pq / N) * 512 + pl * 16, e);
                        T val = e[0];
#pragma unroll
                        for (int j = 1; j < N; ++j) if ((pq % N) == j) val = e[j];
                        F[0][IN][q] = val;
                        const int pmcol = (x0 + pl * V) & ~3;
                        const unsigned w = lds32(ring_s + rofs + ROW_BYTES + pl * 4);
                        kb = (w >> (8 * (x0 + pos - pmcol))) & 1u;
                    }
                    mb = (mb & ~(1u << (V - 1 - q))) | (kb << (V - 1 - q));
                }
            }
        }
        mhist = (MH)(mhist << V) | (MH)mb;
        if (!TMA) { wofs = pofs; pofs = rofs; }    // one wrapping counter serves reads and refills
        rofs = (rofs + SLOT == D * SLOT) ? 0u : rofs + SLOT;
        if (TMA) {
            ridx = (ridx + 1 == D) ? 0u : ridx + 1;
            if (ridx == 0) rphase ^= 1u;
            __syncwarp();                          // every lane is done with the slot about to be refilled
        }
        issue(steady, Phase<1>());                 // refills the slot read one step ago

        T res[V];
        // Skewed: stages run last to first, each reads its window and then the stage before it
        // overwrites that window's oldest row (IU) with the row that becomes "newest" next step.
#pragma unroll
        for (int si = 0; si < K; ++si) {
            const int s = SK ? K - 1 - si : si;
            const T(&rin)[V] = F[s][IN];           // newest row of this stage's input field
            T o[V];
            const unsigned kb = (unsigned)(mhist >> ((s * (R + SK) + R) * V)) & VMASK;
            T s4[V];
            sum4_row<T, V>(F[s][IU], F[s][IC], rin, FL[s], FR[s], s4);
            // halos of the newest row, for the next step; issued after the last use of the old
            // ones so the shuffles can land in the same registers
            FL[s] = shfl_up1(rin[V - 1]);
            FR[s] = shfl_down1(rin[0]);
            if (!CC) {
                // centre row F[s][IC] (row y-s-1), above F[s][IU], below rin
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    // explicit fma: every instantiation (any K, V, code path) rounds identically, so a
                    // solve does not depend on how its sweeps were cut into launches or slabs
                    const T c = F[s][IC][v];
                    const T cand = fma(dt, fma(T(-4), c, s4[v]), c);
                    o[v] = ((kb >> (V - 1 - v)) & 1u) ? c : cand;
                }
            } else {
                // rin is row yy = y-2s of sweep s; hN = L f (row yy-1); output = sweep s+1 of row yy-2
#pragma unroll
                for (int v = 0; v < V; ++v) Hh[s][IN][v] = fma(T(-4), F[s][IC][v], s4[v]);
                if (ZERO) {                        // h is zero-padded like f: its row here is y - s(R+SK) - 1
                    const int hr = y - s * (R + SK) - 1;
                    const bool ghost_row = (hr < 0 && a.top_border) || (hr >= a.rows && a.bottom_border);
#pragma unroll
                    for (int v = 0; v < V; ++v)
                        if (ghost_row || (CEDGE && (col0 + v < 0 || col0 + v >= a.cols))) Hh[s][IN][v] = T(0);
                }
                sum4_row<T, V>(Hh[s][IU], Hh[s][IC], Hh[s][IN], HL[s], HR[s], s4);
                HL[s] = shfl_up1(Hh[s][IN][V - 1]);
                HR[s] = shfl_down1(Hh[s][IN][0]);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const T c = F[s][IU][v];
                    const T cand = fma(c3, Hh[s][IC][v], fma(c2, s4[v], c));
                    o[v] = ((kb >> (V - 1 - v)) & 1u) ? c : cand;
                }
            }
            if (NORM) {
                const int yo = y - LAG;
                if (out_lane && yo >= y0 && yo < y1 && yo >= 0 && yo < a.rows) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        if (col0 + v < a.cols) {
                            const T prev = CC ? F[s][IU][v] : F[s][IC][v];
                            const double dd = (double)(T)(o[v] - prev);
                            n_d2 += dd * dd;
                            n_f2 += (double)o[v] * (double)o[v];
                            const double ad = fabs(dd);
                            if (ad == ad) n_mx = fmax(n_mx, ad); else n_nan = 1.0;
                        }
                    }
                }
            }
            if (s < K - 1) {
#pragma unroll
                for (int v = 0; v < V; ++v) F[s + 1][SK ? IU : IN][v] = o[v];
            } else {
#pragma unroll
                for (int v = 0; v < V; ++v) res[v] = o[v];
            }
        }
        // res is sweep K of row y-LAG
        const int yo = y - LAG;
        if (out_lane && (STEADY || (yo >= y0 && yo < y1))) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                C16 t;
                T *e = reinterpret_cast<T *>(&t);
#pragma unroll
                for (int q = 0; q < N; ++q) e[q] = res[c * N + q];
                reinterpret_cast<C16 *>(dptr)[c] = t;
            }
        }
        dptr += a.pitch;
    };

    int y = ybeg;
    // (not for four fused CCST sweeps: ptxas runs out of predicate registers on the second copy of that loop)
    if (HMI_STREAM_PEEL && MODE == 0 && !TMA && !T2 && !NORM && !(CC && K >= 4)) {
        // head: until the first emitted row reaches y0, in whole groups of three steps (the phases stay aligned)
        const int y_store = ybeg + ((y0 + LAG - ybeg + 2) / 3) * 3;
        const int y_fetch = yend - (D - 1);        // steps below this one still have a row to fetch
        for (; y < y_store && y < ystop; y += 3) {
            step(Phase<0>(), Phase<0>(), y);
            if (y + 1 < ystop) step(Phase<1>(), Phase<0>(), y + 1);
            if (y + 2 < ystop) step(Phase<2>(), Phase<0>(), y + 2);
        }
        for (; y + 2 < y_fetch; y += 3) {          // steady state (ystop > y_fetch always)
            step(Phase<0>(), Phase<1>(), y);
            step(Phase<1>(), Phase<1>(), y + 1);
            step(Phase<2>(), Phase<1>(), y + 2);
        }
    }
    for (; y < ystop; y += 3) {
        step(Phase<0>(), Phase<0>(), y);
        if (y + 1 < ystop) step(Phase<1>(), Phase<0>(), y + 1);
        if (y + 2 < ystop) step(Phase<2>(), Phase<0>(), y + 2);
    }
    if (!TMA && !T2) cp_async_wait<0>();
    // T2: every box that was issued has been waited for (boxes start below yend and the loop runs to
    // ystop >= yend), so no bulk copy is in flight when the warp leaves
}

template <int A, int B> struct MaxOf { static constexpr int value = A > B ? A : B; };

template <typename T, int V, int K, int METHOD, int D, int STG>
struct WarpRing {
    typedef Geo<T, V, K, METHOD> G;
    static constexpr int CP = D * G::SLOT;                                             // cp.async ring (also the edge strips of STG 2)
    static constexpr int value = STG == 1 ? ((D * G::SLOT_T + D * 8 + 15) / 16) * 16
                                          : (STG == 2 ? MaxOf<((CP + 127) / 128) * 128, G::RING_T2>::value : CP);
};

template <typename T, int V, int K, int METHOD, int D, int WPB, int MINB, bool NORM, int STG, int BRD>
__global__ void __launch_bounds__(WPB * 32, MINB) stream_kernel(const __grid_constant__ StreamArgs<T> a)
{
    typedef Geo<T, V, K, METHOD> G;
    constexpr int WARP_RING = WarpRing<T, V, K, METHOD, D, STG>::value;
    constexpr int STG_EDGE = STG == 2 ? 0 : STG;       // strips at a grid edge keep the per-lane copies
    extern __shared__ __align__(128) unsigned char ring_all[];
    __shared__ double sh[NORM ? 128 : 1];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    // tensor-map boxes must land on 128-byte boundaries, which a dynamic shared-memory base does not
    // guarantee (the launcher reserves 128 spare bytes for STG 2)
    const unsigned ring_0 = STG == 2 ? ((smem_u32(ring_all) + 127u) & ~127u) : smem_u32(ring_all);
    const unsigned ring_s = ring_0 + (unsigned)wib * WARP_RING;
    const long long gw = (long long)blockIdx.x * WPB + wib;
    double n_d2 = 0.0, n_f2 = 0.0, n_mx = 0.0, n_nan = 0.0;

    // warp -> (strip, chunk).  The n_edge column-edge strips (first, last) run the slower general
    // path, so they come first and in shorter chunks; the other strips share one chunk height.
    int win, chunk, ch_rows;
    bool valid;
    const long long n_e = (long long)a.n_edge * a.nchunks_e;
    if (gw < n_e) {
        win = (int)(gw % a.n_edge) == 0 ? 0 : a.nwin - 1;
        chunk = (int)(gw / a.n_edge);
        ch_rows = a.chunk_rows_e;
        valid = true;
    } else {
        const long long g = gw - n_e;
        const int ni = a.nwin - a.n_edge;
        valid = ni > 0 && g < (long long)ni * a.nchunks;
        win = ni > 0 ? (a.n_edge ? 1 : 0) + (int)(g % ni) : 0;
        chunk = ni > 0 ? (int)(g / ni) : 0;
        ch_rows = a.chunk_rows;
    }

    if (valid) {
        const int x0 = win * G::S - G::HV;
        const int y0 = a.row_lo + chunk * ch_rows;
        const int y1 = min(a.row_hi, y0 + ch_rows);
        const int ybeg = y0 - G::H, yend = y1 + G::H;
        const bool col_edge = (x0 < 0) || (x0 + 32 * V > a.cols);
        const bool row_edge = (ybeg < 0 && a.top_border) || (yend > a.rows && a.bottom_border) ||
                              (ybeg < a.read_lo) || (yend > a.read_hi);
        if (col_edge)
            stream_rows<T, V, K, METHOD, D, 2, NORM, STG_EDGE, BRD>(a, ring_s, lane, win, y0, y1, n_d2, n_f2, n_mx, n_nan);
        else if (row_edge)
            stream_rows<T, V, K, METHOD, D, 1, NORM, STG_EDGE, BRD>(a, ring_s, lane, win, y0, y1, n_d2, n_f2, n_mx, n_nan);
        else
            stream_rows<T, V, K, METHOD, D, 0, NORM, STG, BRD>(a, ring_s, lane, win, y0, y1, n_d2, n_f2, n_mx, n_nan);
    }
    if (NORM) block_norm_reduce(n_d2, n_f2, n_mx, n_nan, sh, a.partials, a.counter, a.result);
}

// ------------------------------------------------------------------------------------------------
// host side

template <typename T, int V, int K, int METHOD, bool NORM, int STG, int BRD = HMI_BORDER_REFLECT101>
static cudaError_t launch_one(const StreamArgs<T> &a0, cudaStream_t st, int *nblocks_out)
{
    typedef Geo<T, V, K, METHOD> G;
    constexpr int D = HMI_STREAM_RING, WPB = HMI_STREAM_WPB, MINB = HMI_STREAM_MINB;
    constexpr size_t SMEM = (size_t)WPB * WarpRing<T, V, K, METHOD, D, STG>::value + (STG == 2 ? 128 : 0);
    StreamArgs<T> a = a0;
    a.nwin = (a.cols + G::S - 1) / G::S;
    const int nrows = a.row_hi - a.row_lo;
    auto kern = stream_kernel<T, V, K, METHOD, D, WPB, MINB, NORM, STG, BRD>;
    static int cta_slots = 0;         // per instantiation: CTAs resident on the whole GPU at once
    if (cta_slots == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e != cudaSuccess) return e;
        cta_slots = stream_cta_slots((const void *)kern, WPB * 32, SMEM);
    }
    a.n_edge = 0;
    a.nchunks_e = 0;
    a.chunk_rows_e = 0;
    if (a.chunk_rows == 0) {
        a.n_edge = a.nwin >= 2 ? 2 : 1;
        stream_pick_chunks(nrows, a.nwin, a.n_edge, G::H + G::SK * (K - 1), WPB, cta_slots, 1.7, &a.chunk_rows,
                           &a.chunk_rows_e);
        a.nchunks_e = (nrows + a.chunk_rows_e - 1) / a.chunk_rows_e;
    } else if (a.chunk_rows < 0) {
        a.chunk_rows = stream_auto_chunk_rows(nrows, a.nwin, G::H);
    }
    a.nchunks = (nrows + a.chunk_rows - 1) / a.chunk_rows;
    const long long warps = (long long)a.n_edge * a.nchunks_e + (long long)(a.nwin - a.n_edge) * a.nchunks;
    const int blocks = (int)((warps + WPB - 1) / WPB);
    if (nblocks_out) *nblocks_out = blocks;
    static int debug = -1;
    if (debug < 0) debug = getenv("HMI_DEBUG") ? 1 : 0;
    if (debug)
        fprintf(stderr, "[hmi] stream T=%d V=%d K=%d method=%d norm=%d: rows=%d nwin=%d chunk=%d x%d edge=%d x%d "
                        "blocks=%d slots=%d (%.2f waves)\n", (int)sizeof(T), V, K, METHOD, (int)NORM, nrows, a.nwin,
                a.chunk_rows, a.nchunks, a.chunk_rows_e, a.nchunks_e, blocks, cta_slots, (double)blocks / cta_slots);
    kern<<<blocks, WPB * 32, SMEM, st>>>(a);
    return cudaGetLastError();
}

int stream_cta_slots(const void *kernel, int threads, size_t smem)
{
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ < 1) occ = 1;
    cudaGetLastError();
    return sms * occ;
}

void stream_pick_chunks(int nrows, int nwin, int n_edge, int halo, int wpb, int cta_slots, double edge_cost,
                        int *chunk_rows, int *chunk_rows_edge, int min_chunk)
{
    // Wave-aware chunk heights.  Every warp streams (chunk + 2*halo) rows plus a pipeline fill and
    // the launch takes ceil(CTAs / resident CTAs) rounds of that; a fixed chunk height leaves the
    // last round partly empty (4096^2, fp64 CCST: SMs idle for 22 % of the kernel,
    // profiles/r01_ncu_full_ccst_f64_k3_summary.txt).  Pick the chunk count that minimises
    // rounds * rows-per-warp.  The column-edge strips run the general code path (`edge_cost` times
    // the instructions per row: 1.7 here, 2 - 2.8 in hmi_stream_nl.cu), so they get chunks that much
    // shorter and are scheduled first; without that a single full wave is slower than three ragged
    // ones, because the whole launch waits for the two edge strips (profiles/r02_tune_skew_ab.log,
    // profiles/r02_tune_tv_chunks.log).
    const int fill = 8;
    // Load balance.  A launch of r rounds does not take r times one round: all warps of a round start together, the
    // round ends with its slowest warp (edge strips, DRAM page conflicts) and the SMs idle while it drains, so a
    // launch of few long rounds loses to one of a few more, shorter ones although those recompute more overlap rows.
    // Forcing the number of rounds (profiles/r05_tune_picker_min_rounds.log): a 4096 x 32768 fp32 slab, harmonic K = 8,
    // 2 / 3 / 4 / 6 / 8 rounds 2207 / 2400 / 2434 / 2306 / 2175 GLUPS; a 512-row band of the pipelined copies, CCST fp64,
    // 2 / 3 / 4 / 6 rounds 480 / 551 / 548 / 490; 4096^2 CCST fp64 1 / 2 / 3 rounds 597 / 612 / 569.  Half a round of
    // drain time per launch reproduces every one of those optima (and leaves many-wave launches alone).
    static double round_bias = -1.0;
    if (round_bias < 0.0) round_bias = getenv("HMI_PICK_ROUND_BIAS") ? atof(getenv("HMI_PICK_ROUND_BIAS")) : 0.5;
    // (short chunks never pay — a chunk's fixed cost is more than the 2*halo + fill rows counted here: with the
    // load-balance term and no floor the model cut a 4096^2 fp32 TV grid into three rounds of 32-row chunks, 4 % slower
    // than one round of 100-row chunks, and with a 48-row floor into two rounds of 50; the fp32 TV / AMLE kernel, whose
    // rows are the cheapest, asks for 96.  profiles/r05_tune_picker_round_bias_nl_f32.log, ..._min48_summary.txt)
    int lo = 4 * halo;
    if (lo < min_chunk) lo = min_chunk;
    if (lo > nrows) lo = nrows;
    double best = 1e300;
    int best_ch = nrows, best_che = nrows;
    const int max_nch = nrows / lo > 0 ? nrows / lo : 1;
    for (int nch = 1; nch <= max_nch; ++nch) {
        const int ch = (nrows + nch - 1) / nch;
        if (ch < lo) break;
        const int n = (nrows + ch - 1) / ch;                      // chunks this height really gives
        int che = (int)((ch + 2 * halo + fill) / edge_cost) - 2 * halo - fill;
        if (che < 8) che = 8;
        if (che > ch) che = ch;
        const int ne = (nrows + che - 1) / che;
        const long long warps = (long long)n_edge * ne + (long long)(nwin - n_edge) * n;
        const long long ctas = (warps + wpb - 1) / wpb;
        const long long rounds = (ctas + cta_slots - 1) / cta_slots;
        const double cost = ((double)rounds + round_bias) * (ch + 2 * halo + fill) * (1.0 + ch / 8192.0);
        if (cost <= best) { best = cost; best_ch = ch; best_che = che; }
    }
    *chunk_rows = best_ch < 8 ? 8 : best_ch;
    *chunk_rows_edge = best_che < 8 ? 8 : best_che;
}

int stream_auto_chunk_rows(int nrows, int nwin, int halo)
{
    // Legacy rule (chunk_rows < 0), kept for A/B measurements: ~4 waves of 16 warps per SM, within
    // [max(64, 10*halo), 192] rows.
    const long long target_warps = 148LL * 16 * 4;
    long long ch = ((long long)nrows * nwin + target_warps - 1) / target_warps;
    long long lo = 10LL * halo;
    if (lo < 64) lo = 64;
    if (ch < lo) ch = lo;
    if (ch > 192) ch = 192;
    if (ch > nrows) ch = nrows;
    if (ch < 1) ch = 1;
    return (int)ch;
}

int stream_halo(int method, int k) { return k * method_radius(method); }

int stream_cells_per_lane(int dtype, int vec) { return dtype == HMI_F32 ? 4 : (vec == 4 ? 4 : 2); }

int stream_max_k(int method, int dtype)
{
    (void)dtype;
    return method == HMI_CCST ? STREAM_MAX_K_CCST : STREAM_MAX_K_HARMONIC;
}

int stream_max_blocks(int method, int rows_total, int cols)
{
    // upper bound over every (V, K, chunk_rows) this file can launch: strips produce >= 32 columns,
    // chunks are >= 8 rows (callers clamp chunk_rows >= 8)
    const long long nwin = (cols + 31) / 32 + 3;
    const long long nch = (rows_total + 7) / 8 + 1;
    (void)method;
    return (int)((nwin * nch + 3) / 4 + 1);
}

// zero padding (the SciPy convolvers): default cells per lane, cp.async staging
template <typename T, int V>
static cudaError_t launch_zero(int method, int k, const StreamArgs<T> &a, cudaStream_t st, int *nblocks_out)
{
    constexpr int Z = HMI_BORDER_ZERO;
    if (a.do_norm) {
        if (k != 1) return cudaErrorInvalidValue;
        if (method == HMI_HARMONIC) return launch_one<T, V, 1, HMI_HARMONIC, true, 0, Z>(a, st, nblocks_out);
        if (method == HMI_CCST) return launch_one<T, V, 1, HMI_CCST, true, 0, Z>(a, st, nblocks_out);
        return cudaErrorInvalidValue;
    }
#define HMI_CASE(M, KK) \
    if (method == M && k == KK) return launch_one<T, V, KK, M, false, 0, Z>(a, st, nblocks_out);
    HMI_CASE(HMI_HARMONIC, 1) HMI_CASE(HMI_HARMONIC, 2) HMI_CASE(HMI_HARMONIC, 3) HMI_CASE(HMI_HARMONIC, 4)
    HMI_CASE(HMI_HARMONIC, 6)
    HMI_CASE(HMI_CCST, 1) HMI_CASE(HMI_CCST, 2) HMI_CASE(HMI_CCST, 3)
#undef HMI_CASE
    return cudaErrorInvalidValue;
}

bool stream_zero_supports(int method, int k)
{
    return (method == HMI_HARMONIC && ((k >= 1 && k <= 4) || k == 6)) || (method == HMI_CCST && k >= 1 && k <= 3);
}

template <typename T, int V, int TMA>
static cudaError_t launch_v(int method, int k, const StreamArgs<T> &a, cudaStream_t st, int *nblocks_out)
{
    if (a.do_norm) {   // the convergence sums are fused into single-sweep launches only
        if (k != 1) return cudaErrorInvalidValue;
        if (method == HMI_HARMONIC) return launch_one<T, V, 1, HMI_HARMONIC, true, TMA>(a, st, nblocks_out);
        if (method == HMI_CCST) return launch_one<T, V, 1, HMI_CCST, true, TMA>(a, st, nblocks_out);
        return cudaErrorInvalidValue;
    }
#define HMI_CASE(M, KK) \
    if (method == M && k == KK) return launch_one<T, V, KK, M, false, TMA>(a, st, nblocks_out);
#if defined(HMI_STREAM_ONLY_M)   // bisecting builds: one instantiation
    HMI_CASE(HMI_STREAM_ONLY_M, HMI_STREAM_ONLY_K)
#elif defined(HMI_STREAM_LEAN)      // experiment builds (tools/variants): the bench instantiations only
    HMI_CASE(HMI_HARMONIC, 1) HMI_CASE(HMI_HARMONIC, 6) HMI_CASE(HMI_CCST, 1) HMI_CASE(HMI_CCST, 3)
#else
    HMI_CASE(HMI_HARMONIC, 1) HMI_CASE(HMI_HARMONIC, 2) HMI_CASE(HMI_HARMONIC, 3)
    HMI_CASE(HMI_HARMONIC, 4) HMI_CASE(HMI_HARMONIC, 5) HMI_CASE(HMI_HARMONIC, 6)
    HMI_CASE(HMI_HARMONIC, 7) HMI_CASE(HMI_HARMONIC, 8)
    HMI_CASE(HMI_CCST, 1) HMI_CASE(HMI_CCST, 2) HMI_CASE(HMI_CCST, 3) HMI_CASE(HMI_CCST, 4)
#endif
#undef HMI_CASE
    return cudaErrorInvalidValue;
}

template <>
cudaError_t launch_stream<float>(int method, int k, const StreamArgs<float> &a, cudaStream_t st, int *nblocks_out)
{
    if (a.border == HMI_BORDER_ZERO) return launch_zero<float, 4>(method, k, a, st, nblocks_out);
    if (a.tma == 3 && !a.do_norm && pipe_supports(method, k)) return launch_pipe<float>(method, k, a, st, nblocks_out);
#ifndef HMI_STREAM_LEAN
    if (a.tma == 2) return launch_v<float, 4, 2>(method, k, a, st, nblocks_out);
    if (a.tma == 1) return launch_v<float, 4, 1>(method, k, a, st, nblocks_out);
#endif
    return launch_v<float, 4, 0>(method, k, a, st, nblocks_out);
}

template <>
cudaError_t launch_stream<double>(int method, int k, const StreamArgs<double> &a, cudaStream_t st, int *nblocks_out)
{
    if (a.border == HMI_BORDER_ZERO) return launch_zero<double, 2>(method, k, a, st, nblocks_out);
    if (a.tma == 3 && !a.do_norm && pipe_supports(method, k)) return launch_pipe<double>(method, k, a, st, nblocks_out);
    if (a.vec == 4) return launch_v<double, 4, 0>(method, k, a, st, nblocks_out);
#ifndef HMI_STREAM_LEAN
    if (a.tma == 2) return launch_v<double, 2, 2>(method, k, a, st, nblocks_out);
    if (a.tma == 1) return launch_v<double, 2, 1>(method, k, a, st, nblocks_out);
#endif
    return launch_v<double, 2, 0>(method, k, a, st, nblocks_out);
}

}  // namespace hmi
