// Streaming single-sweep kernel for the two non-linear, one-sided-difference methods: total
// variation (tv_inpainter.py:20-33 + differential.py:146-187) and AMLE (amle_inpainter.py:41-59),
// under the reflect-101 border rule.
//
// Same skeleton as hmi_stream.cu (one warp per strip, cp.async shared-memory ring, x-neighbours by
// warp shuffles, y-neighbours in registers, interior/edge code paths, fused convergence sums), with
// what the asymmetric schemes need on top:
//   * Orientation by addressing.  The convolution-oriented scheme (convolver "masked*") is the
//     point mirror of the correlation-oriented one (convolver "opencv"): the kernel always
//     evaluates the s = +1 formulas of SURVEY.md A.3 in *canonical* coordinates and, for s = -1,
//     walks the rows bottom-up and the columns right-to-left (canonical j' = pitch-1-j, i' = rows-1-i).
//   * Intermediate fields with their own border rule.  The reference filters every intermediate
//     array (n = g/|g| for TV, ux/uy for AMLE) with reflect-101 applied to *that array*, which
//     is not what a mirrored ghost of f gives for one-sided differences.  So the row of
//     intermediates is a pipeline stage of its own (row y-1 is built when row y lands, row y-2
//     is emitted), and the two places where the rule bites are overridden explicitly: the left
//     neighbour of the first column is the value at column 1, the upper neighbour of the first
//     row is the value at row 1.
//   * One reciprocal square root per cell (rsqrt + multiplies) instead of the reference's sqrt and
//     two divides: within a few ulp, far inside the 1e-10 / 1e-5 parity bars, and 3x fewer
//     slow-path fp64 operations than the generic kernel, which evaluates the normalisation at the
//     three cells a step touches.
#include <cstdio>
#include <cstdlib>

#include "hmi_stream_nl.cuh"
#include "hmi_stream.cuh"
#include "hmi_stream_common.cuh"

namespace hmi {

template <typename T, int V>
struct GeoNL {
    static constexpr int N = Chunk16<T>::N;
    static constexpr int NCH = V / N;
    static constexpr int HV = V;                        // one lane of halo per side
    static constexpr int S = 32 * V - 2 * HV;
    static constexpr int ROW_BYTES = 32 * V * (int)sizeof(T);
    static constexpr int SLOT = ROW_BYTES + 128;
    static_assert(V % N == 0 && NCH >= 1, "a lane moves whole 16-byte chunks");
    static_assert(V == 2 || V == 4, "a lane's mask bytes must sit inside one aligned 4-byte word");
};

// fp32: the bare hardware approximation (MUFU.RSQ, 2 ulp) — rsqrtf() wraps it in a denormal-input
// rescaling that a sum of squares plus a positive constant never needs.
__device__ __forceinline__ float rsqrt_t(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// fp64: the hardware seed (MUFU.RSQ64H, ~2^-21) and one cubic refinement step, without the special-case call
// CUDA's rsqrt() carries: the argument is a sum of squares plus a positive constant, always normal.
__device__ __forceinline__ double rsqrt_t(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    // one cubically convergent step: e = 1 - x y^2 (|e| < 2^-20), y <- y + y e (1/2 + 3/8 e); error ~ e^3
    const double e = fma(-x, y * y, 1.0);
    return fma(y * e, fma(e, 0.375, 0.5), y);
}

// known ? c : cand.  With three fused fp32 stages the 12 predicate-based selects of a row step exhaust
// ptxas's 7 predicate registers ("register allocation failed"), so that variant blends by bit mask.
template <int K>
__device__ __forceinline__ float select_known(unsigned known, float c, float cand)
{
    if (K < 3) return known ? c : cand;
    const unsigned m = 0u - known;
    return __uint_as_float((__float_as_uint(c) & m) | (__float_as_uint(cand) & ~m));
}
template <int K>
__device__ __forceinline__ double select_known(unsigned known, double c, double cand) { return known ? c : cand; }

// One warp, one strip, rows [y0, y1) of it, in canonical coordinates (see header).
// MODE 0: interior (lean loop); 1: only the rows need care (reflect-101 at the global top/bottom,
// clamping to the slab's memory, the "upper neighbour of row 0" rule); 2: ghost columns as well.
// The row loop is unrolled by three so that the windows of f rows, lane halos and intermediate rows
// rotate by register renaming: slot(row) = row mod 3.
// K sweeps are fused as K chained stages (temporal blocking): stage s takes the row stage s-1 has just
// emitted as its newest input row, so step y emits sweep K of row y - 2K and only that is stored.
// One sweep consumes one column and one row of validity per side; the strip already carries a whole
// lane (V cells) of halo per side, so K <= V costs no extra column overlap.
// BRD = border rule every filtered array (f of every sweep and the intermediates) carries: reflect-101
// ("opencv", "masked*"), zero padding (the SciPy convolvers, convolver.py:17-18, 47-54) or SciPy's symmetric
// 'reflect' (AMLE convolve_in_1d, amle_inpainter.py:26-37).  It only changes what a ghost cell holds.
template <typename T, int V, int METHOD, int SGN, int D, int MODE, bool NORM, int K, int BRD>
__device__ __forceinline__ void stream_rows_nl(const StreamNLArgs<T> &a, const unsigned ring_s, const int lane,
                                               const int win, const int y0, const int y1, double &n_d2,
                                               double &n_f2, double &n_mx, double &n_nan)
{
    typedef GeoNL<T, V> G;
    typedef typename Chunk16<T>::type C16;
    constexpr int N = G::N, NCH = G::NCH, HV = G::HV, S = G::S, ROW_BYTES = G::ROW_BYTES, SLOT = G::SLOT;
    constexpr bool TV = (METHOD == HMI_TV);
    constexpr bool REDGE = MODE >= 1, CEDGE = MODE == 2;
    constexpr bool SYM = (BRD == HMI_BORDER_SYMMETRIC), ZERO = (BRD == HMI_BORDER_ZERO);
    constexpr unsigned VMASK = (1u << V) - 1u;
    constexpr int LAG = 2 * K;                         // step y emits sweep K of row y - 2K
    static_assert(K >= 1 && K <= 3 && K <= V, "temporal block deeper than the lane halo");
    static_assert(!NORM || K == 1, "the convergence sums are fused into single-sweep launches only");
    static_assert((LAG + 1) * V <= 32, "mask history does not fit 32 bits");

    const int x0 = a.cbase + win * S - HV;             // canonical column of lane 0, cell 0
    const int col0 = x0 + lane * V;                    // canonical column of this lane's cell 0
    const int acol = SGN > 0 ? col0 : (int)a.pitch - col0 - V;   // actual column of the lane's chunk
    const int ybeg = y0 - K, yend = y1 + K;            // canonical rows fetched: [ybeg, yend)
    const int ystop = y1 + LAG;                        // steps: [ybeg, ystop)
    const bool out_lane = (lane >= 1) && (lane < 31) && (col0 < a.chi) && (col0 + V > a.clo);
    const T dt = a.dt;
    const int rtop = a.top_border ? 0 : -0x40000000;   // canonical row whose upper neighbour is row 1
    const int rbot = a.bottom_border ? a.rows : 0x40000000;   // first ghost row below the grid

    const int mcol = acol & ~3;
    const unsigned msh = (unsigned)(acol & 3) * 8u;
    const bool f_copies = CEDGE ? ((acol >= 0) && (acol + V <= (int)a.pitch)) : true;
    const bool m_copies = CEDGE ? ((mcol >= 0) && (mcol + 4 <= (int)a.mpitch)) : true;

    const unsigned ring_f = ring_s + lane * 16;
    const unsigned ring_m = ring_s + ROW_BYTES + lane * 4;
    const T *gbase = a.src + acol;
    const uint8_t *mbase = a.mask + mcol;
    const long long rstep = SGN > 0 ? a.pitch : -a.pitch;        // actual-row stride of a canonical step
    const long long mstep = SGN > 0 ? a.mpitch : -a.mpitch;
    auto arow = [&](int yc) { return SGN > 0 ? yc : a.rows - 1 - yc; };

    int ynext = ybeg;
    unsigned wofs = 0;
    const T *gnext = gbase + (long long)arow(ybeg) * a.pitch;
    const uint8_t *mnext = mbase + (long long)arow(ybeg) * a.mpitch;

    auto issue = [&]() {
        if (ynext < yend) {
            const T *prow = gnext;
            const uint8_t *mrow = mnext;
            if (REDGE) {                               // reflect-101 at the global edges, clamp to memory
                int ry = ynext;
                if (ZERO) ;                            // ghost rows are zeroed when they are consumed
                else if (ry < 0 && a.top_border) ry = SYM ? -ry - 1 : -ry;
                else if (ry >= a.rows && a.bottom_border) ry = SYM ? 2 * a.rows - 1 - ry : 2 * (a.rows - 1) - ry;
                ry = max(a.read_lo, min(a.read_hi - 1, ry));
                prow = gbase + (long long)arow(ry) * a.pitch;
                mrow = mbase + (long long)arow(ry) * a.mpitch;
            }
            if (f_copies) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) cp_async16(ring_f + wofs + c * 512, prow + c * N);
            }
            if (m_copies) cp_async4(ring_m + wofs, mrow);
        }
        cp_async_commit();
        ++ynext;
        if (!REDGE) { gnext += rstep; mnext += mstep; }
        wofs = (wofs + SLOT == D * SLOT) ? 0u : wofs + SLOT;
    };

#pragma unroll
    for (int u = 0; u < D - 1; ++u) issue();

    // per stage, slot(row) = row mod 3:  its input rows ys-2, ys-1, ys (and ys-3 until row ys lands in
    // its slot), their lane halos, and the intermediate rows ys-3, ys-2, ys-1 (n = g/|g| for TV,
    // ux/uy for AMLE); ys = y - 2s is the newest input row of stage s at step y
    T F[K][3][V], P[K][3][V], Q[K][3][V], HL[K][3], HR[K][3];
#pragma unroll
    for (int s = 0; s < K; ++s)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            HL[s][q] = HR[s][q] = T(0);
#pragma unroll
            for (int v = 0; v < V; ++v) F[s][q][v] = P[s][q][v] = Q[s][q][v] = T(0);
        }
    unsigned mhist = 0u;
    unsigned rofs = 0;
    T *dptr = a.dst + (long long)arow(ybeg - LAG) * a.pitch + acol;

    auto step = [&](auto ph, const int y) {
        constexpr int PH = decltype(ph)::value;
        constexpr int IU = PH, IC = (PH + 1) % 3, IN = (PH + 2) % 3;   // rows ys-2, ys-1, ys (and ys-3)
        T res[V];
#pragma unroll
        for (int s = 0; s < K; ++s) {
            const int ys = y - 2 * s;
            // AMLE: w0 = f[r+1] - f[r-1] of the row emitted below, r = ys-2; row ys-3 still sits in slot IN
            T v0[V];
            if (!TV) {
#pragma unroll
                for (int v = 0; v < V; ++v) v0[v] = F[s][IC][v] - F[s][IN][v];
            }
            if (s == 0) {
                cp_async_wait<D - 2>();
                if (CEDGE) __syncwarp();
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    T e[N];
                    lds16(ring_f + rofs + c * 512, e);
#pragma unroll
                    for (int q = 0; q < N; ++q) {
                        const int act = c * N + q;     // actual element -> canonical cell
                        F[0][IN][SGN > 0 ? act : V - 1 - act] = e[q];
                    }
                }
                unsigned mb;                           // canonical cell v -> bit V-1-v
                {
                    const unsigned w = lds32(ring_m + rofs);
                    if (V == 4) mb = (w * (SGN > 0 ? 0x08040201u : 0x01020408u)) >> 24;
                    else mb = ((((w >> msh) & 0xffffu) * (SGN > 0 ? 0x0201u : 0x0102u)) >> 8) & 3u;
                }
                if (CEDGE) {
                    mb &= VMASK;
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const int c = col0 + v;
                        if (ZERO) {
                            if (c < a.clo || c >= a.chi) F[0][IN][v] = T(0);
                        } else if (c < a.clo || c >= a.chi) { // ghost column: value and mask of the mirror cell
                            const int cm = SYM ? ((c < a.clo) ? 2 * a.clo - c - 1 : 2 * a.chi - 1 - c)
                                               : ((c < a.clo) ? 2 * a.clo - c : 2 * (a.chi - 1) - c);
                            const int pos = max(0, min(32 * V - 1, cm - x0));
                            const int pl = pos / V, pv = pos % V;
                            const int pact = SGN > 0 ? pv : V - 1 - pv;    // actual element in lane pl
                            T e[N];
                            lds16(ring_s + rofs + (pact / N) * 512 + pl * 16, e);
                            T val = e[0];
#pragma unroll
                            for (int j = 1; j < N; ++j) if ((pact % N) == j) val = e[j];
                            F[0][IN][v] = val;
                            const int pacol = SGN > 0 ? x0 + pl * V : (int)a.pitch - (x0 + pl * V) - V;
                            const int pcol = SGN > 0 ? x0 + pos : (int)a.pitch - 1 - (x0 + pos);
                            const unsigned w = lds32(ring_s + rofs + ROW_BYTES + pl * 4);
                            const unsigned kb = (w >> (8 * (pcol - (pacol & ~3)))) & 1u;
                            mb = (mb & ~(1u << (V - 1 - v))) | (kb << (V - 1 - v));
                        }
                    }
                }
                if (ZERO && REDGE) {
                    if (ys < rtop || ys >= rbot) {     // a row outside the grid
#pragma unroll
                        for (int v = 0; v < V; ++v) F[0][IN][v] = T(0);
                    }
                }
                mhist = (mhist << V) | mb;
                rofs = (rofs + SLOT == D * SLOT) ? 0u : rofs + SLOT;
                issue();
            } else {
                // newest input row of this stage = the row the previous stage has just emitted.  The
                // reference re-applies reflect-101 to every sweep's field, and for these one-sided
                // schemes a ghost cell does not evolve like its mirror image, so ghosts are re-imposed:
#pragma unroll
                for (int v = 0; v < V; ++v) F[s][IN][v] = res[v];
                if (REDGE) {
                    if (ZERO) {
                        if (ys < rtop || ys >= rbot) { // rows outside the grid hold zeros
#pragma unroll
                            for (int v = 0; v < V; ++v) F[s][IN][v] = T(0);
                        }
                    } else if (ys == rbot) {           // first row below the grid := row rows-2 (slot IU); symmetric: rows-1 (IC)
#pragma unroll
                        for (int v = 0; v < V; ++v) F[s][IN][v] = SYM ? F[s][IC][v] : F[s][IU][v];
                    }
                }
                if (CEDGE && ZERO) {
#pragma unroll
                    for (int v = 0; v < V; ++v)
                        if (col0 + v < a.clo || col0 + v >= a.chi) F[s][IN][v] = T(0);
                } else if (CEDGE) {                    // ghost columns := their mirror columns (other lanes)
                    T fix[V];
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const int c = col0 + v;
                        const bool ghost = (c < a.clo || c >= a.chi);
                        const int cm = SYM ? ((c < a.clo) ? 2 * a.clo - c - 1 : 2 * a.chi - 1 - c)
                                           : ((c < a.clo) ? 2 * a.clo - c : 2 * (a.chi - 1) - c);
                        const int pos = max(0, min(32 * V - 1, cm - x0));
                        const int pl = pos / V, pv = pos % V;
                        T val = F[s][IN][v];
#pragma unroll
                        for (int j = 0; j < V; ++j) {
                            const T cand = __shfl_sync(0xffffffffu, F[s][IN][j], pl);
                            if (ghost && pv == j) val = cand;
                        }
                        fix[v] = val;
                    }
#pragma unroll
                    for (int v = 0; v < V; ++v) F[s][IN][v] = fix[v];
                }
            }

            HL[s][IN] = shfl_up1(F[s][IN][V - 1]);
            HR[s][IN] = shfl_down1(F[s][IN][0]);

            // ---- intermediates of row ys-1 (centre F[IC], row below = F[IN]) -> P[IC], Q[IC]
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T right = (v < V - 1) ? F[s][IC][v + 1] : HR[s][IC];
                const T dx = right - F[s][IC][v];      // f[i, j+1] - f[i, j]
                const T dy = F[s][IN][v] - F[s][IC][v];   // f[i+1, j] - f[i, j]
                if (TV) {
                    const T inv = rsqrt_t(fma(dx, dx, fma(dy, dy, a.eps2)));
                    P[s][IC][v] = dx * inv;            // n_x
                    Q[s][IC][v] = dy * inv;            // n_y
                } else {
                    P[s][IC][v] = dy;                  // ux (reference's "x" is the row direction)
                    Q[s][IC][v] = dx;                  // uy
                }
            }

            // ---- emit row r = ys-2 (centre F[IU]); its intermediates are P[IU], Q[IU], the row above
            // them (ys-3) sits in slot IN
            const int r = ys - 2;
            if (REDGE) {
                if (r == rtop) {                       // the intermediate above row 0: reflect-101 [-1] := [1], symmetric := [0], zero := 0
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        P[s][IN][v] = ZERO ? T(0) : (SYM ? P[s][IU][v] : P[s][IC][v]);
                        Q[s][IN][v] = ZERO ? T(0) : (SYM ? Q[s][IU][v] : Q[s][IC][v]);
                    }
                    if (!TV && s > 0 && !ZERO) {       // ... and the field itself: f[-1] := f[1], so w0 = 0 (symmetric: f[-1] := f[0])
#pragma unroll
                        for (int v = 0; v < V; ++v) v0[v] = SYM ? F[s][IC][v] - F[s][IU][v] : T(0);
                    }
                }
            }
            const T Pl = shfl_up1(P[s][IU][V - 1]);
            const T Ql = TV ? T(0) : shfl_up1(Q[s][IU][V - 1]);
            T Pr = T(0), Qr = T(0);
            if (CEDGE) { Pr = shfl_down1(P[s][IU][0]); if (!TV) Qr = shfl_down1(Q[s][IU][0]); }
            const unsigned kb = (mhist >> ((2 * s + 2) * V)) & VMASK;
            T out[V];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                T pl = (v > 0) ? P[s][IU][v - 1] : Pl;
                T ql = (v > 0) ? Q[s][IU][v - 1] : Ql;
                if (CEDGE && col0 + v == a.clo) {      // the intermediate left of column 0: reflect-101 [-1] := [1], symmetric := [0], zero := 0
                    if (ZERO) { pl = T(0); ql = T(0); }
                    else if (SYM) { pl = P[s][IU][v]; ql = Q[s][IU][v]; }
                    else {
                        pl = (v < V - 1) ? P[s][IU][v + 1] : Pr;
                        ql = (v < V - 1) ? Q[s][IU][v + 1] : Qr;
                    }
                }
                T stp;
                if (TV) {
                    stp = (P[s][IU][v] - pl) + (Q[s][IU][v] - Q[s][IN][v]);
                } else {
                    const T uxx = P[s][IU][v] - P[s][IN][v], uxy = P[s][IU][v] - pl;
                    const T uyx = Q[s][IU][v] - Q[s][IN][v], uyy = Q[s][IU][v] - ql;
                    const T fl = (v > 0) ? F[s][IU][v - 1] : HL[s][IU];
                    const T fr = (v < V - 1) ? F[s][IU][v + 1] : HR[s][IU];
                    T w0 = v0[v];
                    T w1 = fr - fl;
                    const T inv = rsqrt_t(fma(w0, w0, fma(w1, w1, T(1e-15))));
                    w0 *= inv;
                    w1 *= inv;
                    stp = (uxx * w0) * w0 + (uyy * w1) * w1 + ((uxy + uyx) * w0) * w1;
                }
                const T c = F[s][IU][v];
                const T cand = c + dt * stp;
                out[v] = select_known<K>((kb >> (V - 1 - v)) & 1u, c, cand);
            }
            if (NORM) {
                if (out_lane && r >= y0 && r < y1 && r >= 0 && r < a.rows) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        if (col0 + v >= a.clo && col0 + v < a.chi) {
                            const double dd = (double)(T)(out[v] - F[s][IU][v]);
                            n_d2 += dd * dd;
                            n_f2 += (double)out[v] * (double)out[v];
                            const double ad = fabs(dd);
                            if (ad == ad) n_mx = fmax(n_mx, ad); else n_nan = 1.0;
                        }
                    }
                }
            }
#pragma unroll
            for (int v = 0; v < V; ++v) res[v] = out[v];
        }
        const int r = y - LAG;                         // res is sweep K of row y - 2K
        if (out_lane && r >= y0 && r < y1) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                C16 t;
                T *e = reinterpret_cast<T *>(&t);
#pragma unroll
                for (int q = 0; q < N; ++q) {
                    const int act = c * N + q;
                    e[q] = res[SGN > 0 ? act : V - 1 - act];
                }
                reinterpret_cast<C16 *>(dptr)[c] = t;
            }
        }
        dptr += rstep;
    };

    for (int y = ybeg; y < ystop; y += 3) {
        step(Phase<0>(), y);
        if (y + 1 < ystop) step(Phase<1>(), y + 1);
        if (y + 2 < ystop) step(Phase<2>(), y + 2);
    }
    cp_async_wait<0>();
}

// Register budget: ptxas left to itself settles on 80 registers for TV and spills the loop bounds to
// local memory (an LDL on the loop's critical path, profiles/r02_ncu_full_tv_f64_v2_summary.txt);
// asking for MINB resident CTAs instead caps it at 65536 / (128 MINB) registers without spills.
template <typename T, int METHOD, int K> struct NLBudget {
    static constexpr int MINB = K > 2 ? 2 : (K > 1 ? 3 : ((METHOD == HMI_TV) ? 5 : 4));
};

template <typename T, int V, int METHOD, int SGN, int D, int WPB, bool NORM, int K, int BRD>
__global__ void __launch_bounds__(WPB * 32, NLBudget<T, METHOD, K>::MINB) stream_nl_kernel(const StreamNLArgs<T> a)
{
    typedef GeoNL<T, V> G;
    extern __shared__ __align__(16) unsigned char ring_all[];
    __shared__ double sh[NORM ? 128 : 1];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const unsigned ring_s = smem_u32(ring_all) + (unsigned)wib * D * G::SLOT;
    const long long gw = (long long)blockIdx.x * WPB + wib;
    double n_d2 = 0.0, n_f2 = 0.0, n_mx = 0.0, n_nan = 0.0;

    // warp -> (strip, chunk): the n_edge column-edge strips (first, last) run the general path, so
    // they are scheduled first and in shorter chunks (see stream_pick_chunks in hmi_stream.cu)
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
        const int x0 = a.cbase + win * G::S - G::HV;
        const int y0 = a.row_lo + chunk * ch_rows;
        const int y1 = min(a.row_hi, y0 + ch_rows);
        const bool col_edge = (x0 < a.clo) || (x0 + 32 * V > a.chi);
        const bool row_edge = (y0 - K < 0 && a.top_border) || (y1 + K > a.rows && a.bottom_border) ||
                              (y0 - K < a.read_lo) || (y1 + K > a.read_hi) || (y0 <= 0 && a.top_border);
        if (col_edge)
            stream_rows_nl<T, V, METHOD, SGN, D, 2, NORM, K, BRD>(a, ring_s, lane, win, y0, y1, n_d2, n_f2, n_mx, n_nan);
        else if (row_edge)
            stream_rows_nl<T, V, METHOD, SGN, D, 1, NORM, K, BRD>(a, ring_s, lane, win, y0, y1, n_d2, n_f2, n_mx, n_nan);
        else
            stream_rows_nl<T, V, METHOD, SGN, D, 0, NORM, K, BRD>(a, ring_s, lane, win, y0, y1, n_d2, n_f2, n_mx, n_nan);
    }
    if (NORM) block_norm_reduce(n_d2, n_f2, n_mx, n_nan, sh, a.partials, a.counter, a.result);
}

// ------------------------------------------------------------------------------------------------
// host side

template <typename T, int V, int METHOD, int SGN, bool NORM, int K, int BRD>
static cudaError_t launch_nl_one(const NLLaunch<T> &l, cudaStream_t st, int *nblocks_out)
{
    typedef GeoNL<T, V> G;
    constexpr int D = 8, WPB = 4;
    constexpr size_t SMEM = (size_t)WPB * D * G::SLOT;
    StreamNLArgs<T> a;
    a.src = l.src; a.dst = l.dst; a.mask = l.mask;
    a.pitch = l.pitch; a.mpitch = l.mpitch;
    a.rows = l.rows;
    a.dt = l.dt; a.eps2 = l.eps2;
    a.partials = l.partials; a.counter = l.counter; a.result = l.result;
    // canonical frame: identity for s = +1, point mirror (rows-1-i, pitch-1-j) for s = -1
    if (SGN > 0) {
        a.clo = 0; a.chi = l.cols;
        a.row_lo = l.row_lo; a.row_hi = l.row_hi;
        a.read_lo = l.read_lo; a.read_hi = l.read_hi;
        a.top_border = l.top_border; a.bottom_border = l.bottom_border;
    } else {
        a.clo = (int)l.pitch - l.cols; a.chi = (int)l.pitch;
        a.row_lo = l.rows - l.row_hi; a.row_hi = l.rows - l.row_lo;
        a.read_lo = l.rows - l.read_hi; a.read_hi = l.rows - l.read_lo;
        a.top_border = l.bottom_border; a.bottom_border = l.top_border;
    }
    a.cbase = a.clo - (a.clo % V);
    a.nwin = (a.chi - a.cbase + G::S - 1) / G::S;
    const int nrows = a.row_hi - a.row_lo;
    auto kern = stream_nl_kernel<T, V, METHOD, SGN, D, WPB, NORM, K, BRD>;
    static int cta_slots = 0;
    if (cta_slots == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e != cudaSuccess) return e;
        cta_slots = stream_cta_slots((const void *)kern, WPB * 32, SMEM);
    }
    a.n_edge = 0;
    a.nchunks_e = 0;
    a.chunk_rows_e = 0;
    if (l.chunk_rows == 0) {
        a.n_edge = a.nwin >= 2 ? 2 : 1;
        // general path / lean path, instructions per row (SASS): fp32 2.8, fp64 2.0; be generous — the edge
        // strips are 2 of ~70 and come first
        stream_pick_chunks(nrows, a.nwin, a.n_edge, K, WPB, cta_slots, V == 4 ? 3.5 : 2.5, &a.chunk_rows,
                           &a.chunk_rows_e, V == 4 ? 96 : 48);
        a.nchunks_e = (nrows + a.chunk_rows_e - 1) / a.chunk_rows_e;
    } else {
        a.chunk_rows = l.chunk_rows > 0 ? l.chunk_rows : stream_nl_auto_chunk_rows(nrows, a.nwin);
    }
    a.nchunks = (nrows + a.chunk_rows - 1) / a.chunk_rows;
    const long long warps = (long long)a.n_edge * a.nchunks_e + (long long)(a.nwin - a.n_edge) * a.nchunks;
    const int blocks = (int)((warps + WPB - 1) / WPB);
    if (nblocks_out) *nblocks_out = blocks;
    static int debug = -1;
    if (debug < 0) debug = getenv("HMI_DEBUG") ? 1 : 0;
    if (debug)
        fprintf(stderr, "[hmi] stream_nl T=%d V=%d K=%d method=%d sgn=%d norm=%d: rows=%d nwin=%d chunk=%d x%d edge=%d x%d "
                        "blocks=%d slots=%d (%.2f waves)\n", (int)sizeof(T), V, K, METHOD, SGN, (int)NORM, nrows, a.nwin,
                a.chunk_rows, a.nchunks, a.chunk_rows_e, a.nchunks_e, blocks, cta_slots, (double)blocks / cta_slots);
    kern<<<blocks, WPB * 32, SMEM, st>>>(a);
    return cudaGetLastError();
}

int stream_nl_max_k(int method, int dtype, int border)
{
    return (method == HMI_TV && dtype == HMI_F32 && border == HMI_BORDER_REFLECT101) ? 3 : 2;
}

bool stream_nl_supports(int method, int sgn, int border)
{
    if (method != HMI_TV && method != HMI_AMLE) return false;
    if (border == HMI_BORDER_REFLECT101) return true;
    if (border == HMI_BORDER_ZERO) return sgn == -1;
    return border == HMI_BORDER_SYMMETRIC && method == HMI_AMLE && sgn == -1;
}

int stream_nl_auto_chunk_rows(int nrows, int nwin)
{
    // same rule as stream_auto_chunk_rows with a one-row halo: ~4 waves, within [32, 192] rows
    const long long target_warps = 148LL * 16 * 4;
    long long ch = ((long long)nrows * nwin + target_warps - 1) / target_warps;
    if (ch < 32) ch = 32;
    if (ch > 192) ch = 192;
    if (ch > nrows) ch = nrows;
    if (ch < 1) ch = 1;
    return (int)ch;
}

// three fused sweeps need three cells of lane halo: fp32 (V = 4) only; and TV only — ptxas cannot
// allocate the predicates of the three-stage AMLE loop
template <typename T, int V, int METHOD, int SGN, int BRD>
static cudaError_t launch_nl_k3(const NLLaunch<T> &l, cudaStream_t st, int *nb)
{
    if constexpr (V >= 3 && METHOD == HMI_TV && BRD == HMI_BORDER_REFLECT101) return launch_nl_one<T, V, METHOD, SGN, false, 3, BRD>(l, st, nb);
    else return cudaErrorInvalidValue;
}

template <typename T, int V>
static cudaError_t launch_nl_v(const NLLaunch<T> &l, cudaStream_t st, int *nb)
{
    if (l.do_norm && l.k != 1) return cudaErrorInvalidValue;   // the norm is fused into single-sweep launches
#define HMI_NL(M, SG, B)                                                               \
    if (l.method == M && l.sgn == SG && l.border == B) {                               \
        if (l.do_norm) return launch_nl_one<T, V, M, SG, true, 1, B>(l, st, nb);       \
        if (l.k == 1) return launch_nl_one<T, V, M, SG, false, 1, B>(l, st, nb);       \
        if (l.k == 2) return launch_nl_one<T, V, M, SG, false, 2, B>(l, st, nb);       \
        if (l.k == 3) return launch_nl_k3<T, V, M, SG, B>(l, st, nb);                  \
        return cudaErrorInvalidValue;                                                  \
    }
    HMI_NL(HMI_TV, 1, HMI_BORDER_REFLECT101) HMI_NL(HMI_TV, -1, HMI_BORDER_REFLECT101)
    HMI_NL(HMI_AMLE, 1, HMI_BORDER_REFLECT101) HMI_NL(HMI_AMLE, -1, HMI_BORDER_REFLECT101)
    // the SciPy convolvers are convolutions (s = -1) with zero padding; AMLE convolve_in_1d is s = -1, symmetric
    HMI_NL(HMI_TV, -1, HMI_BORDER_ZERO) HMI_NL(HMI_AMLE, -1, HMI_BORDER_ZERO)
    HMI_NL(HMI_AMLE, -1, HMI_BORDER_SYMMETRIC)
#undef HMI_NL
    return cudaErrorInvalidValue;
}

template <>
cudaError_t launch_stream_nl<float>(const NLLaunch<float> &l, cudaStream_t st, int *nb)
{
    return launch_nl_v<float, 4>(l, st, nb);
}
template <>
cudaError_t launch_stream_nl<double>(const NLLaunch<double> &l, cudaStream_t st, int *nb)
{
    return launch_nl_v<double, 2>(l, st, nb);
}

}  // namespace hmi
