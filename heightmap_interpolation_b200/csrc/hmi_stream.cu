// Streaming temporal-blocked sweep kernel for the symmetric stencils (harmonic, CCST/biharmonic)
// under the reflect-101 border rule — the hot kernel of the path.
//
// Design (B200-first, no counterpart in the reference, which runs one full-grid NumPy/OpenCV pass
// per operator: fd_pde_inpainter.py:311-326):
//   * One WARP owns a column strip of 32*V cells and streams down a chunk of rows.  Each lane holds
//     V consecutive cells of the current row in registers (one or two 16-byte loads per lane,
//     coalesced across the warp); x-neighbours come from warp shuffles, y-neighbours from the
//     previous rows kept in registers.  No shared memory and no block barrier on the hot loop.
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
#include "hmi_common.cuh"
#include "hmi_stream.cuh"

namespace hmi {

template <typename T> struct Chunk16;
template <> struct Chunk16<float> { typedef float4 type; static constexpr int N = 4; };
template <> struct Chunk16<double> { typedef double2 type; static constexpr int N = 2; };

template <typename T, int V>
__device__ __forceinline__ void load_row_vec(const T *__restrict__ p, T (&v)[V])
{
    typedef typename Chunk16<T>::type C;
    constexpr int N = Chunk16<T>::N;
#pragma unroll
    for (int c = 0; c < V / N; ++c) {
        const C t = __ldg(reinterpret_cast<const C *>(p) + c);
        const T *e = reinterpret_cast<const T *>(&t);
#pragma unroll
        for (int q = 0; q < N; ++q) v[c * N + q] = e[q];
    }
}

template <typename T, int V>
__device__ __forceinline__ void store_row_vec(T *__restrict__ p, const T (&v)[V])
{
    typedef typename Chunk16<T>::type C;
    constexpr int N = Chunk16<T>::N;
#pragma unroll
    for (int c = 0; c < V / N; ++c) {
        C t;
        T *e = reinterpret_cast<T *>(&t);
#pragma unroll
        for (int q = 0; q < N; ++q) e[q] = v[c * N + q];
        reinterpret_cast<C *>(p)[c] = t;
    }
}

template <int V>
__device__ __forceinline__ unsigned load_mask_vec(const uint8_t *__restrict__ p)
{
    // returns V bits, bit v set = cell v is KNOWN
    unsigned bits = 0;
    if (V == 4) {
        const uchar4 m = __ldg(reinterpret_cast<const uchar4 *>(p));
        bits = (m.x ? 1u : 0u) | (m.y ? 2u : 0u) | (m.z ? 4u : 0u) | (m.w ? 8u : 0u);
    } else if (V == 2) {
        const uchar2 m = __ldg(reinterpret_cast<const uchar2 *>(p));
        bits = (m.x ? 1u : 0u) | (m.y ? 2u : 0u);
    } else {
#pragma unroll
        for (int v = 0; v < V; ++v) bits |= (__ldg(p + v) ? 1u : 0u) << v;
    }
    return bits;
}

// One loaded row: V values + V mask bits
template <typename T, int V>
struct Row {
    T v[V];
    unsigned m;
};

template <typename T, int V>
__device__ __forceinline__ void load_row(const StreamArgs<T> &a, int y, int col0, bool edge, Row<T, V> &r)
{
    // row index with the reflect-101 rule at global edges, clamped to the rows that exist
    int ry = y;
    if (ry < 0 && a.top_border) ry = -ry;
    else if (ry >= a.rows && a.bottom_border) ry = 2 * (a.rows - 1) - ry;
    ry = max(a.read_lo, min(a.read_hi - 1, ry));
    const T *prow = a.src + (long long)ry * a.pitch;
    const uint8_t *mrow = a.mask + (long long)ry * a.mpitch;
    if (!edge) {
        load_row_vec<T, V>(prow + col0, r.v);
        r.m = load_mask_vec<V>(mrow + col0);
    } else {
        unsigned bits = 0;
#pragma unroll
        for (int q = 0; q < V; ++q) {
            int c = col0 + q;
            if (c < 0) c = -c;
            else if (c >= a.cols) c = 2 * (a.cols - 1) - c;
            c = max(0, min(a.cols - 1, c));
            r.v[q] = __ldg(prow + c);
            bits |= (__ldg(mrow + c) ? 1u : 0u) << q;
        }
        r.m = bits;
    }
}

template <typename T>
__device__ __forceinline__ T shfl_up1(T v) { return __shfl_up_sync(0xffffffffu, v, 1); }
template <typename T>
__device__ __forceinline__ T shfl_down1(T v) { return __shfl_down_sync(0xffffffffu, v, 1); }

// 5-point Laplacian of the centre row b given the rows above/below and b's lane halos
template <typename T, int V>
__device__ __forceinline__ void lap_row(const T (&up)[V], const T (&b)[V], const T (&dn)[V], T bl, T br,
                                        T (&out)[V])
{
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const T l = (v > 0) ? b[v - 1] : bl;
        const T r = (v < V - 1) ? b[v + 1] : br;
        out[v] = ((up[v] + l) + (r + dn[v])) - T(4) * b[v];
    }
}

template <typename T, int V, int K, int METHOD, int PF, int WPB, int MINB>
__global__ void __launch_bounds__(WPB * 32, MINB) stream_kernel(const StreamArgs<T> a)
{
    constexpr int R = (METHOD == HMI_CCST) ? 2 : 1;
    constexpr int H = K * R;                       // rows/columns of halo one launch consumes
    constexpr int HV = ((H + V - 1) / V) * V;      // column halo rounded to whole lanes
    constexpr int S = 32 * V - 2 * HV;             // columns a warp produces
    constexpr unsigned VM = (1u << V) - 1u;
    static_assert(S > 0, "temporal block too deep for this strip width");

    __shared__ double sh[128];
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * WPB + (threadIdx.x >> 5);
    const int win = (int)(gw % a.nwin);
    const int chunk = (int)(gw / a.nwin);
    double n_d2 = 0.0, n_f2 = 0.0, n_mx = 0.0, n_nan = 0.0;

    if (chunk < a.nchunks) {
        const int x0 = win * S - HV;
        const int col0 = x0 + lane * V;
        const bool edge = (x0 < 0) || (x0 + 32 * V > a.cols);
        const int y0 = a.row_lo + chunk * a.chunk_rows;
        const int y1 = min(a.row_hi, y0 + a.chunk_rows);
        const bool out_lane = (lane >= HV / V) && (lane < 32 - HV / V) && (col0 < a.cols);
        const T dt = a.dt;

        // pipeline state
        T FA[K][V], FB[K][V], FBL[K], FBR[K];
        T HA[K][V], HB[K][V], HBL[K], HBR[K];      // CCST only (dead code for harmonic)
        unsigned long long mhist = 0ull;
#pragma unroll
        for (int s = 0; s < K; ++s) {
            FBL[s] = FBR[s] = HBL[s] = HBR[s] = T(0);
#pragma unroll
            for (int v = 0; v < V; ++v) { FA[s][v] = FB[s][v] = HA[s][v] = HB[s][v] = T(0); }
        }

        const int ybeg = y0 - H, yend = y1 + H;    // rows streamed: [ybeg, yend)
        Row<T, V> pre[PF];
#pragma unroll
        for (int u = 0; u < PF; ++u)
            if (ybeg + u < yend) load_row<T, V>(a, ybeg + u, col0, edge, pre[u]);

        for (int yb = ybeg; yb < yend; yb += PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int y = yb + u;
                if (y < yend) {
                    T cur[V];
#pragma unroll
                    for (int v = 0; v < V; ++v) cur[v] = pre[u].v[v];
                    mhist = (mhist << V) | (unsigned long long)pre[u].m;
                    if (y + PF < yend) load_row<T, V>(a, y + PF, col0, edge, pre[u]);

#pragma unroll
                    for (int s = 0; s < K; ++s) {
                        T out[V];
                        if (METHOD == HMI_HARMONIC) {
                            // centre row FB[s] (row y-s-1), above FA[s], below cur
                            const unsigned kb = (unsigned)(mhist >> ((s + 1) * V)) & VM;
                            T lp[V];
                            lap_row<T, V>(FA[s], FB[s], cur, FBL[s], FBR[s], lp);
#pragma unroll
                            for (int v = 0; v < V; ++v) {
                                const T c = FB[s][v];
                                const T cand = c + dt * lp[v];
                                out[v] = ((kb >> v) & 1u) ? c : cand;
                            }
                            if (s == K - 1 && a.do_norm) {
                                const int yo = y - H;
                                if (out_lane && yo >= y0 && yo < y1 && yo >= 0 && yo < a.rows) {
#pragma unroll
                                    for (int v = 0; v < V; ++v) {
                                        if (col0 + v < a.cols) {
                                            const T d = out[v] - FB[s][v];
                                            const double dd = (double)d;
                                            n_d2 += dd * dd;
                                            n_f2 += (double)out[v] * (double)out[v];
                                            const double ad = fabs(dd);
                                            if (ad == ad) n_mx = fmax(n_mx, ad); else n_nan = 1.0;
                                        }
                                    }
                                }
                            }
                            // shift the window
#pragma unroll
                            for (int v = 0; v < V; ++v) { FA[s][v] = FB[s][v]; FB[s][v] = cur[v]; }
                            FBL[s] = shfl_up1(cur[V - 1]);
                            FBR[s] = shfl_down1(cur[0]);
                        } else {
                            // CCST macro-stage: cur is row yy = y-2s of sweep s
                            //   hC = L f (row yy-1), b = L h (row yy-2), out = sweep s+1 of row yy-2
                            const unsigned kb = (unsigned)(mhist >> ((2 * s + 2) * V)) & VM;
                            T hC[V], bh[V];
                            lap_row<T, V>(FA[s], FB[s], cur, FBL[s], FBR[s], hC);
                            lap_row<T, V>(HA[s], HB[s], hC, HBL[s], HBR[s], bh);
#pragma unroll
                            for (int v = 0; v < V; ++v) {
                                const T c = FA[s][v];
                                const T step = a.tension * HB[s][v] - a.one_minus_tension * bh[v];
                                const T cand = c + dt * step;
                                out[v] = ((kb >> v) & 1u) ? c : cand;
                            }
                            if (s == K - 1 && a.do_norm) {
                                const int yo = y - H;
                                if (out_lane && yo >= y0 && yo < y1 && yo >= 0 && yo < a.rows) {
#pragma unroll
                                    for (int v = 0; v < V; ++v) {
                                        if (col0 + v < a.cols) {
                                            const T d = out[v] - FA[s][v];
                                            const double dd = (double)d;
                                            n_d2 += dd * dd;
                                            n_f2 += (double)out[v] * (double)out[v];
                                            const double ad = fabs(dd);
                                            if (ad == ad) n_mx = fmax(n_mx, ad); else n_nan = 1.0;
                                        }
                                    }
                                }
                            }
#pragma unroll
                            for (int v = 0; v < V; ++v) {
                                FA[s][v] = FB[s][v]; FB[s][v] = cur[v];
                                HA[s][v] = HB[s][v]; HB[s][v] = hC[v];
                            }
                            FBL[s] = shfl_up1(cur[V - 1]);
                            FBR[s] = shfl_down1(cur[0]);
                            HBL[s] = shfl_up1(hC[V - 1]);
                            HBR[s] = shfl_down1(hC[0]);
                        }
#pragma unroll
                        for (int v = 0; v < V; ++v) cur[v] = out[v];
                    }
                    // cur is now sweep K of row y-H
                    const int yo = y - H;
                    if (out_lane && yo >= y0 && yo < y1)
                        store_row_vec<T, V>(a.dst + (long long)yo * a.pitch + col0, cur);
                }
            }
        }
    }
    if (a.do_norm) block_norm_reduce(n_d2, n_f2, n_mx, n_nan, sh, a.partials, a.counter, a.result);
}

// ------------------------------------------------------------------------------------------------
// host side

template <typename T, int V, int K, int METHOD>
static cudaError_t launch_one(const StreamArgs<T> &a0, cudaStream_t st, int *nblocks_out)
{
    constexpr int R = (METHOD == HMI_CCST) ? 2 : 1;
    constexpr int H = K * R, HV = ((H + V - 1) / V) * V, S = 32 * V - 2 * HV;
    constexpr int PF = 4, WPB = 4, MINB = 2;
    StreamArgs<T> a = a0;
    a.nwin = (a.cols + S - 1) / S;
    const int nrows = a.row_hi - a.row_lo;
    if (a.chunk_rows <= 0) a.chunk_rows = stream_auto_chunk_rows(nrows, a.nwin, H);
    a.nchunks = (nrows + a.chunk_rows - 1) / a.chunk_rows;
    const long long warps = (long long)a.nwin * a.nchunks;
    const int blocks = (int)((warps + WPB - 1) / WPB);
    if (nblocks_out) *nblocks_out = blocks;
    stream_kernel<T, V, K, METHOD, PF, WPB, MINB><<<blocks, WPB * 32, 0, st>>>(a);
    return cudaGetLastError();
}

int stream_auto_chunk_rows(int nrows, int nwin, int halo)
{
    // enough warps for ~2 waves of 12 warps on each of the 148 SMs, but chunks tall enough that the
    // 2*halo redundant rows stay a small fraction
    const long long target_warps = 148LL * 12 * 2;
    long long chunks = (target_warps + nwin - 1) / nwin;
    if (chunks < 1) chunks = 1;
    long long ch = (nrows + chunks - 1) / chunks;
    const long long min_ch = 12LL * halo;
    if (ch < min_ch) ch = min_ch;
    if (ch > nrows) ch = nrows;
    if (ch < 1) ch = 1;
    return (int)ch;
}

int stream_halo(int method, int k) { return k * method_radius(method); }

int stream_max_k(int method, int dtype)
{
    (void)dtype;
    return method == HMI_CCST ? STREAM_MAX_K_CCST : STREAM_MAX_K_HARMONIC;
}

int stream_max_blocks(int method, int rows_total, int cols)
{
    // upper bound over every (K, chunk_rows >= 1) this file can launch: strips are >= 32 columns
    // wide after the overlap, chunks at least 1 row; callers clamp chunk_rows >= 8
    const long long nwin = (cols + 31) / 32 + 1;
    const long long nch = (rows_total + 7) / 8 + 1;
    (void)method;
    return (int)((nwin * nch + 3) / 4 + 1);
}

template <typename T>
cudaError_t launch_stream(int method, int k, const StreamArgs<T> &a, cudaStream_t st, int *nblocks_out)
{
    constexpr int V = 4;
#define HMI_CASE(M, KK) \
    if (method == M && k == KK) return launch_one<T, V, KK, M>(a, st, nblocks_out);
    HMI_CASE(HMI_HARMONIC, 1) HMI_CASE(HMI_HARMONIC, 2) HMI_CASE(HMI_HARMONIC, 3)
    HMI_CASE(HMI_HARMONIC, 4) HMI_CASE(HMI_HARMONIC, 5) HMI_CASE(HMI_HARMONIC, 6)
    HMI_CASE(HMI_HARMONIC, 7) HMI_CASE(HMI_HARMONIC, 8)
    HMI_CASE(HMI_CCST, 1) HMI_CASE(HMI_CCST, 2) HMI_CASE(HMI_CCST, 3) HMI_CASE(HMI_CCST, 4)
#undef HMI_CASE
    return cudaErrorInvalidValue;
}

template cudaError_t launch_stream<float>(int, int, const StreamArgs<float> &, cudaStream_t, int *);
template cudaError_t launch_stream<double>(int, int, const StreamArgs<double> &, cudaStream_t, int *);

}  // namespace hmi
