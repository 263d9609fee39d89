// Warp-specialised temporal pipeline for the symmetric stencils (harmonic, CCST/biharmonic) under
// the reflect-101 border rule: the K fused relaxation sweeps of a launch run on K different warps.
//
// Why (profiles/r04_ncu_ccst_f64_k3_32768*_summary.txt): the one-warp-per-strip kernel (hmi_stream.cu)
// keeps all K pipeline stages of a strip in one warp's registers.  With 2 fp64 cells per lane it is
// issue bound — an fp64 instruction holds the issue port for two cycles and 65 % of the instruction
// stream is shuffles, selects, moves and addressing — and with 4 cells per lane the K = 3 windows need
// 234 registers, 8 warps per SM, and the kernel turns latency bound at the same speed.  Here:
//   * a CTA owns one column strip of 32*V cells and one chunk of rows, and is K + 1 warps: warp 0
//     loads, warp 1 + s computes sweep s + 1 (fd_pde_inpainter.py:311-326, one loop iteration per
//     stage).  A stage warp holds one 3-row window of its input (and, CCST, of h = L f), so V = 4 / 8
//     cells per lane fit in ~100 / ~170 registers and the x-neighbour traffic per cell halves again;
//   * rows travel between warps through shared-memory rings in batches of three (the row loop is
//     unrolled by three so the windows rotate by renaming), guarded by mbarriers (full / empty per
//     slot).  The loader fills its ring with cp.async 16-byte copies several batches ahead and
//     publishes a batch when its copies have landed; the left / right halo cells of a lane are read
//     from the neighbouring lanes' part of the row in shared memory, not shuffled;
//   * ring rows are stored in planes of 16-byte chunks (chunk c of lane l at c*512 + l*16), so every
//     128-bit shared-memory access of a warp is conflict free;
//   * the mask row travels once, into a longer ring every stage reads its bits from;
//   * reflect-101: the loader fetches ghost rows from their mirror rows and fills ghost columns from their
//     mirror cells (value and mask) inside the landed row before it publishes the batch.
//     The stencils are symmetric and the sums are associated (up + down) + (left + right), so a ghost
//     cell evolves bit for bit like its mirror image through the later stages, which therefore know
//     nothing about borders (same argument as in hmi_stream.cu; results are bit-identical to it).
// Only the last stage writes to global memory: HBM traffic per cell update is ~17/K (fp64), 9/K (fp32) B.
#include <cstdio>
#include <cstdlib>

#include "hmi_common.cuh"
#include "hmi_stream.cuh"
#include "hmi_stream_common.cuh"

#ifndef HMI_PIPE_NS
#define HMI_PIPE_NS 3
#endif
#ifndef HMI_PIPE_NSL
#define HMI_PIPE_NSL 4
#endif

namespace hmi {

constexpr int pipe_pow2_at_least(int x) { int p = 1; while (p < x) p *= 2; return p; }

template <typename T, int V, int K, int METHOD>
struct PipeGeo {
    static constexpr int N = Chunk16<T>::N;             // cells per 16-byte chunk
    static constexpr int NCH = V / N;                   // 16-byte chunks per lane per row
    static constexpr int R = (METHOD == HMI_CCST) ? 2 : 1;
    static constexpr int H = K * R;                     // rows/columns of halo one launch consumes
    static constexpr int HV = ((H + V - 1) / V) * V;    // column halo rounded to whole lanes
    static constexpr int S = 32 * V - 2 * HV;           // columns a CTA produces
    static constexpr int ROWB = 32 * V * (int)sizeof(T);
    static constexpr int MROWB = 32 * V;                // mask bytes of one strip row
    static constexpr int MW = V / 4;                    // 4-byte mask words per lane per row
    static constexpr int RB = 3;                        // rows per hand-off batch (= unroll of the row loop)
    static constexpr int NSL = HMI_PIPE_NSL;            // loader ring: batches
    static constexpr int INFL = NSL - 1;                // ... of which this many may be in flight
    static constexpr int NS = HMI_PIPE_NS;              // rings between stages: batches
    // the loader runs at most (NSL + (K-1) NS + 1) batches ahead of the last stage, which reads the
    // mask of the row H behind its input row
    static constexpr int M = pipe_pow2_at_least((NSL + (K - 1) * NS + 2) * RB + H + 2);
    static constexpr int LOADER_RING = NSL * RB * ROWB;
    static constexpr int STAGE_RING = NS * RB * ROWB;
    static constexpr int MASK_RING = M * MROWB;
    static constexpr int NBAR = 2 * NSL + 2 * NS * (K - 1);
    static constexpr int SMEM = LOADER_RING + (K - 1) * STAGE_RING + MASK_RING + ((NBAR * 8 + 127) / 128) * 128;
    static constexpr int THREADS = (K + 1) * 32;
    static_assert(V % N == 0 && V % 4 == 0, "a lane moves whole 16-byte chunks and whole mask words");
    static_assert(S > 0, "temporal block too deep for this strip width");
};

// mbarrier wait with a time-out (about 2 s): a protocol error traps the launch instead of hanging the GPU
__device__ __forceinline__ unsigned mbar_try(unsigned bar, unsigned parity)
{
    unsigned done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done;
}
__device__ __forceinline__ void mbar_wait_b(unsigned bar, unsigned parity)
{
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity))
        if (clock64() - t0 > 4000000000LL) __trap();
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sts16(unsigned addr, const double (&v)[2])
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(addr), "d"(v[0]), "d"(v[1]) : "memory");
}
__device__ __forceinline__ void sts16(unsigned addr, const float (&v)[4])
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])
                 : "memory");
}
__device__ __forceinline__ void sts1(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void sts1(unsigned addr, float v) { asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts8(unsigned addr, unsigned v) { asm volatile("st.shared.u8 [%0], %1;\n" ::"r"(addr), "r"(v) : "memory"); }

// ------------------------------------------------------------------------------------------------
// warp 0: fetch rows [ybeg, yend) of the strip into the loader ring, their mask bytes into the mask ring.
// MODE 0 = interior strip and rows; 1 = rows need reflecting / clamping; 2 = ghost columns as well.
template <typename T, int V, int K, int METHOD, int MODE>
__device__ __forceinline__ void pipe_loader(const StreamArgs<T> &a, const unsigned ringL, const unsigned ringM,
                                            const unsigned barL, const int lane, const int x0, const int ybeg,
                                            const int yend)
{
    typedef PipeGeo<T, V, K, METHOD> G;
    constexpr int N = G::N, NCH = G::NCH, RB = G::RB, NSL = G::NSL, INFL = G::INFL, ROWB = G::ROWB, MROWB = G::MROWB,
                  M = G::M, MW = G::MW;
    const unsigned fullL = barL, emptyL = barL + 8 * NSL;
    const int nrows = yend - ybeg;
    const int nb = (nrows + RB - 1) / RB;
    // chunk g = lane + 32 i of the row goes to compute lane g / NCH, plane g % NCH
    unsigned dofs[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int g = lane + 32 * i;
        dofs[i] = (unsigned)((g % NCH) * 512 + (g / NCH) * 16);
    }
    const T *grow = a.src + (long long)ybeg * a.pitch + x0;            // MODE 0: incremental
    const uint8_t *mrow = a.mask + (long long)ybeg * a.mpitch + x0;

    auto copy_row = [&](const int y, const unsigned dst, const unsigned mdst) {
        const T *prow = grow;
        const uint8_t *pm = mrow;
        if (MODE >= 1) {
            int ry = y;
            if (ry < 0 && a.top_border) ry = -ry;
            else if (ry >= a.rows && a.bottom_border) ry = 2 * (a.rows - 1) - ry;
            ry = max(a.read_lo, min(a.read_hi - 1, ry));
            prow = a.src + (long long)ry * a.pitch + x0;
            pm = a.mask + (long long)ry * a.mpitch + x0;
        }
        // MODE 2: whatever lies inside the allocation is copied as it is (padding columns included);
        // the ghost columns are filled in from their mirror cells when the batch is published
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int g = lane + 32 * i;
            const int c0 = x0 + g * N;
            if (MODE < 2 || (c0 >= 0 && c0 + N <= (int)a.pitch)) cp_async16(dst + dofs[i], prow + g * N);
        }
#pragma unroll
        for (int i = 0; i < MW; ++i) {
            const int w = lane + 32 * i;
            const int c0 = x0 + 4 * w;
            if (MODE < 2 || (c0 >= 0 && c0 + 4 <= (int)a.mpitch)) cp_async4(mdst + 4 * w, pm + 4 * w);
        }
        if (MODE == 0) { grow += a.pitch; mrow += a.mpitch; }
    };
    // ghost columns of a landed row take the value and the mask of their mirror cell, which sits in the
    // same strip row (the strip reaches at least HV >= H columns into the grid)
    auto mirror_row = [&](const unsigned dst, const unsigned mdst) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int g = lane + 32 * i;
#pragma unroll
            for (int q = 0; q < N; ++q) {
                const int c = x0 + g * N + q;
                if (c < 0 || c >= a.cols) {
                    const int cm = c < 0 ? -c : 2 * (a.cols - 1) - c;
                    const int p = max(0, min(32 * V - 1, cm - x0));
                    const int pv = p % V;
                    const T v = lds1(dst + (unsigned)((pv / N) * 512 + (p / V) * 16 + (pv % N) * (int)sizeof(T)), T(0));
                    sts1(dst + dofs[i] + q * (unsigned)sizeof(T), v);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < MW; ++i) {
            const int w = lane + 32 * i;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = x0 + 4 * w + q;
                if (c < 0 || c >= a.cols) {
                    const int cm = c < 0 ? -c : 2 * (a.cols - 1) - c;
                    const int p = max(0, min(32 * V - 1, cm - x0));
                    sts8(mdst + 4 * w + q, lds8(mdst + p));
                }
            }
        }
    };
    auto publish = [&](const int bp) {                 // this lane's copies of batch bp have landed
        const int sl = bp % NSL;
        if (MODE == 2) {
            __syncwarp();                              // ... and the other lanes'
#pragma unroll
            for (int r = 0; r < RB; ++r)
                if (bp * RB + r < nrows)
                    mirror_row(ringL + (unsigned)((sl * RB + r) * ROWB), ringM + (unsigned)(((bp * RB + r) & (M - 1)) * MROWB));
        }
        mbar_arrive(fullL + 8 * sl);
    };

    int slot = 0, wraps = 0;
    for (int b = 0; b < nb; ++b) {
        if (wraps > 0) mbar_wait_b(emptyL + 8 * slot, (unsigned)(wraps - 1) & 1u);    // stage 0 is done with the batch that sat here
        const int j0 = b * RB;
#pragma unroll
        for (int r = 0; r < RB; ++r)
            if (j0 + r < nrows)
                copy_row(ybeg + j0 + r, ringL + (unsigned)((slot * RB + r) * ROWB), ringM + (unsigned)(((j0 + r) & (M - 1)) * MROWB));
        cp_async_commit();
        if (b >= INFL) {
            cp_async_wait<INFL>();                     // batch b - INFL has landed
            publish(b - INFL);
        }
        if (++slot == NSL) { slot = 0; ++wraps; }
    }
    cp_async_wait<0>();
    for (int b = max(nb - INFL, 0); b < nb; ++b) publish(b);
}

// ------------------------------------------------------------------------------------------------
// warp 1 + s: sweep s + 1.  Reads its input rows from `ring_in` (the loader's ring for s = 0), writes
// sweep s + 1 of the row R behind into `ring_out`, or, the last stage, into global memory.
template <typename T, int V, int K, int METHOD, bool LAST>
__device__ __forceinline__ void pipe_stage(const StreamArgs<T> &a, const int s, const unsigned ring_in, const int ns_in,
                                           const unsigned full_in, const unsigned empty_in, const unsigned ring_out,
                                           const unsigned full_out, const unsigned empty_out, const unsigned ringM,
                                           const int lane, const int x0, const int y0, const int y1)
{
    typedef PipeGeo<T, V, K, METHOD> G;
    typedef typename Chunk16<T>::type C16;
    constexpr int N = G::N, NCH = G::NCH, R = G::R, H = G::H, HV = G::HV, RB = G::RB, NS = G::NS, ROWB = G::ROWB,
                  MROWB = G::MROWB, M = G::M, MW = G::MW;
    constexpr bool CC = (METHOD == HMI_CCST);
    const int ybeg = y0 - H, yend = y1 + H;
    const int nrows = yend - ybeg;
    const int nb = (nrows + RB - 1) / RB;
    const int col0 = x0 + lane * V;
    const bool out_lane = (lane >= HV / V) && (lane < 32 - HV / V) && (col0 < a.cols);
    const T dt = a.dt;
    // CCST with folded coefficients: f' = f + c2*(sum of h's four neighbours) + c3*h,
    //   c2 = -dt(1-t), c3 = dt*t - 4*c2      (== f + dt*(t*h - (1-t)*L h)), as in hmi_stream.cu
    const T c2 = -(dt * a.one_minus_tension);
    const T c3 = fma(T(-4), c2, dt * a.tension);

    // 3-row windows, rotating by renaming: in phase PH rows above / centre / new sit in slots PH, PH+1, PH+2 (mod 3)
    T F[3][V], Hh[CC ? 3 : 1][V];
    T FL = T(0), FR = T(0), HL = T(0), HR = T(0);
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int v = 0; v < V; ++v) {
            F[q][v] = T(0);
            if (CC) Hh[q][v] = T(0);
        }
    const unsigned own = (unsigned)lane * 16u;                         // this lane's chunk inside a plane
    const unsigned lft = (unsigned)((NCH - 1) * 512 + max(lane - 1, 0) * 16);   // left neighbour's last chunk
    const unsigned rgt = (unsigned)(min(lane + 1, 31) * 16);                    // right neighbour's first chunk
    const int mrel = R * (s + 1);                                      // this stage emits the row mrel behind its input row
    T *dptr = a.dst + (long long)(ybeg - H) * a.pitch + col0;          // LAST: row emitted by step 0

    int slot_in = 0, par_in = 0, slot_out = 0, wraps_out = 0;

    auto step = [&](auto ph, const int j) {
        constexpr int PH = decltype(ph)::value;
        constexpr int IU = PH, IC = (PH + 1) % 3, IN = (PH + 2) % 3;
        const unsigned row = ring_in + (unsigned)((slot_in * RB + PH) * ROWB);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            T e[N];
            lds16(row + c * 512 + own, e);
#pragma unroll
            for (int q = 0; q < N; ++q) F[IN][c * N + q] = e[q];
        }
        T el[N], er[N];
        lds16(row + lft, el);
        lds16(row + rgt, er);
        // mask bits of the row this step emits: cell v -> bit V-1-v
        const unsigned ma = ringM + (unsigned)(((j - mrel) & (M - 1)) * MROWB) + (unsigned)(lane * V);
        unsigned kb = 0;
#pragma unroll
        for (int i = 0; i < MW; ++i) kb = (kb << 4) | ((lds32(ma + 4 * i) * 0x08040201u) >> 24);

        T s4[V], o[V];
        sum4_row<T, V>(F[IU], F[IC], F[IN], FL, FR, s4);
        FL = el[N - 1];                                               // halos of the newest row, for the next step
        FR = er[0];
        if constexpr (!CC) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T c = F[IC][v];
                const T cand = fma(dt, fma(T(-4), c, s4[v]), c);
                o[v] = ((kb >> (V - 1 - v)) & 1u) ? c : cand;
            }
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v) Hh[IN][v] = fma(T(-4), F[IC][v], s4[v]);
            sum4_row<T, V>(Hh[IU], Hh[IC], Hh[IN], HL, HR, s4);
            HL = shfl_up1(Hh[IN][V - 1]);
            HR = shfl_down1(Hh[IN][0]);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T c = F[IU][v];
                const T cand = fma(c3, Hh[IC][v], fma(c2, s4[v], c));
                o[v] = ((kb >> (V - 1 - v)) & 1u) ? c : cand;
            }
        }
        if (LAST) {
            const int yo = ybeg + j - H;
            if (out_lane && yo >= y0 && yo < y1) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    C16 t;
                    T *e = reinterpret_cast<T *>(&t);
#pragma unroll
                    for (int q = 0; q < N; ++q) e[q] = o[c * N + q];
                    reinterpret_cast<C16 *>(dptr)[c] = t;
                }
            }
            dptr += a.pitch;
        } else {
            const unsigned orow = ring_out + (unsigned)((slot_out * RB + PH) * ROWB) + own;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                T e[N];
#pragma unroll
                for (int q = 0; q < N; ++q) e[q] = o[c * N + q];
                sts16(orow + c * 512, e);
            }
        }
    };

    for (int b = 0; b < nb; ++b) {
        const int j0 = b * RB;
        mbar_wait_b(full_in + 8 * slot_in, (unsigned)par_in);                                   // the batch has arrived
        if (!LAST && wraps_out > 0) mbar_wait_b(empty_out + 8 * slot_out, (unsigned)(wraps_out - 1) & 1u);   // the next stage has let go of the slot
        step(Phase<0>(), j0);
        if (j0 + 1 < nrows) step(Phase<1>(), j0 + 1);
        if (j0 + 2 < nrows) step(Phase<2>(), j0 + 2);
        __syncwarp();                                  // every lane has read the batch / written its part of the new one
        if (lane == 0) {
            mbar_arrive(empty_in + 8 * slot_in);
            if (!LAST) mbar_arrive(full_out + 8 * slot_out);
        }
        if (++slot_in == ns_in) { slot_in = 0; par_in ^= 1; }
        if (!LAST && ++slot_out == NS) { slot_out = 0; ++wraps_out; }
    }
}

template <typename T, int V, int K, int METHOD, int MINB>
__global__ void __launch_bounds__((K + 1) * 32, MINB) pipe_kernel(const __grid_constant__ StreamArgs<T> a)
{
    typedef PipeGeo<T, V, K, METHOD> G;
    extern __shared__ __align__(128) unsigned char pipe_smem[];
    const unsigned base = smem_u32(pipe_smem);
    const unsigned ringL = base;
    const unsigned ringS = base + G::LOADER_RING;
    const unsigned ringM = ringS + (K - 1) * G::STAGE_RING;
    const unsigned bars = ringM + G::MASK_RING;
    const unsigned barL = bars;                                   // fullL[NSL], emptyL[NSL]
    const unsigned barS = bars + 8 * 2 * G::NSL;                  // per boundary: full[NS], empty[NS]
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < G::NSL; ++i) {
            mbar_init(barL + 8 * i, 32);                          // every loader lane publishes its own copies
            mbar_init(barL + 8 * (G::NSL + i), 1);
        }
#pragma unroll
        for (int i = 0; i < 2 * G::NS * (K - 1); ++i) mbar_init(barS + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    // CTA -> (strip, chunk), as in stream_kernel: the column-edge strips come first, in shorter chunks
    const long long gw = blockIdx.x;
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
    if (!valid) return;
    const int x0 = win * G::S - G::HV;
    const int y0 = a.row_lo + chunk * ch_rows;
    const int y1 = min(a.row_hi, y0 + ch_rows);
    const int ybeg = y0 - G::H, yend = y1 + G::H;

    // role of this warp: 0 loads, 1 + s is stage s.  Rotated by the CTA index, because consecutive warps of
    // a CTA sit on consecutive scheduler partitions: without it one of the four partitions of an SM would
    // hold nothing but loaders (K + 1 = 4) and a quarter of the issue slots would idle
    const int role = (wib + (int)(blockIdx.x % (K + 1))) % (K + 1);
    if (role == 0) {
        const bool col_edge = (x0 < 0) || (x0 + 32 * V > a.cols);
        const bool row_edge = (ybeg < 0 && a.top_border) || (yend > a.rows && a.bottom_border) ||
                              (ybeg < a.read_lo) || (yend > a.read_hi);
        if (col_edge) pipe_loader<T, V, K, METHOD, 2>(a, ringL, ringM, barL, lane, x0, ybeg, yend);
        else if (row_edge) pipe_loader<T, V, K, METHOD, 1>(a, ringL, ringM, barL, lane, x0, ybeg, yend);
        else pipe_loader<T, V, K, METHOD, 0>(a, ringL, ringM, barL, lane, x0, ybeg, yend);
    } else {
        const int s = role - 1;
        const unsigned ring_in = s == 0 ? ringL : ringS + (unsigned)((s - 1) * G::STAGE_RING);
        const unsigned full_in = s == 0 ? barL : barS + (unsigned)((s - 1) * 2 * G::NS * 8);
        const unsigned empty_in = s == 0 ? barL + 8 * G::NSL : full_in + 8 * G::NS;
        const int ns_in = s == 0 ? G::NSL : G::NS;
        if (s == K - 1) {
            pipe_stage<T, V, K, METHOD, true>(a, s, ring_in, ns_in, full_in, empty_in, 0u, 0u, 0u, ringM, lane, x0, y0, y1);
        } else {
            const unsigned ring_out = ringS + (unsigned)(s * G::STAGE_RING);
            const unsigned full_out = barS + (unsigned)(s * 2 * G::NS * 8);
            pipe_stage<T, V, K, METHOD, false>(a, s, ring_in, ns_in, full_in, empty_in, ring_out, full_out,
                                               full_out + 8 * G::NS, ringM, lane, x0, y0, y1);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side

template <typename T, int V, int K, int METHOD>
static cudaError_t launch_pipe_one(const StreamArgs<T> &a0, cudaStream_t st, int *nblocks_out)
{
    typedef PipeGeo<T, V, K, METHOD> G;
    constexpr int MINB = 1;
    StreamArgs<T> a = a0;
    a.nwin = (a.cols + G::S - 1) / G::S;
    const int nrows = a.row_hi - a.row_lo;
    auto kern = pipe_kernel<T, V, K, METHOD, MINB>;
    static int cta_slots = 0;         // per instantiation: CTAs resident on the whole GPU at once
    if (cta_slots == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
        if (e != cudaSuccess) return e;
        cta_slots = stream_cta_slots((const void *)kern, G::THREADS, G::SMEM);
    }
    a.n_edge = 0;
    a.nchunks_e = 0;
    a.chunk_rows_e = 0;
    if (a.chunk_rows <= 0) {
        a.n_edge = a.nwin >= 2 ? 2 : 1;
        // the stage warps do not see the borders; only the loader of an edge strip is slower
        stream_pick_chunks(nrows, a.nwin, a.n_edge, G::H + K * G::RB, 1, cta_slots, 1.3, &a.chunk_rows, &a.chunk_rows_e);
        a.nchunks_e = (nrows + a.chunk_rows_e - 1) / a.chunk_rows_e;
    }
    a.nchunks = (nrows + a.chunk_rows - 1) / a.chunk_rows;
    const long long ctas = (long long)a.n_edge * a.nchunks_e + (long long)(a.nwin - a.n_edge) * a.nchunks;
    if (nblocks_out) *nblocks_out = (int)ctas;
    static int debug = -1;
    if (debug < 0) debug = getenv("HMI_DEBUG") ? 1 : 0;
    if (debug)
        fprintf(stderr, "[hmi] pipe T=%d V=%d K=%d method=%d: rows=%d nwin=%d chunk=%d x%d edge=%d x%d ctas=%lld slots=%d "
                        "(%.2f waves) smem=%d\n", (int)sizeof(T), V, K, METHOD, nrows, a.nwin, a.chunk_rows, a.nchunks,
                a.chunk_rows_e, a.nchunks_e, ctas, cta_slots, (double)ctas / cta_slots, (int)G::SMEM);
    kern<<<(unsigned)ctas, G::THREADS, G::SMEM, st>>>(a);
    return cudaGetLastError();
}

template <typename T, int V>
static cudaError_t launch_pipe_v(int method, int k, const StreamArgs<T> &a, cudaStream_t st, int *nblocks_out)
{
#define HMI_CASE(M, KK) \
    if (method == M && k == KK) return launch_pipe_one<T, V, KK, M>(a, st, nblocks_out);
    HMI_CASE(HMI_HARMONIC, 2) HMI_CASE(HMI_HARMONIC, 3) HMI_CASE(HMI_HARMONIC, 4) HMI_CASE(HMI_HARMONIC, 5)
    HMI_CASE(HMI_HARMONIC, 6) HMI_CASE(HMI_HARMONIC, 7) HMI_CASE(HMI_HARMONIC, 8)
    HMI_CASE(HMI_CCST, 2) HMI_CASE(HMI_CCST, 3) HMI_CASE(HMI_CCST, 4)
#undef HMI_CASE
    return cudaErrorInvalidValue;
}

bool pipe_supports(int method, int k)
{
    return (method == HMI_HARMONIC && k >= 2 && k <= 8) || (method == HMI_CCST && k >= 2 && k <= 4);
}

template <>
cudaError_t launch_pipe<float>(int method, int k, const StreamArgs<float> &a, cudaStream_t st, int *nblocks_out)
{
    return launch_pipe_v<float, 8>(method, k, a, st, nblocks_out);
}

template <>
cudaError_t launch_pipe<double>(int method, int k, const StreamArgs<double> &a, cudaStream_t st, int *nblocks_out)
{
    if (a.vec == 8) return launch_pipe_v<double, 8>(method, k, a, st, nblocks_out);
    return launch_pipe_v<double, 4>(method, k, a, st, nblocks_out);
}

}  // namespace hmi
