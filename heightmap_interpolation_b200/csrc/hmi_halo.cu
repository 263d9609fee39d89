// Peer-memory halo links between row slabs (include/hmi_b200.h: hmi_halo_*, hmi_sweeps_exchange).
//
// The reference is one process holding the whole grid (fd_pde_inpainter.py:281-366); cut into row
// slabs, the only data a sweep needs from another GPU are the `ghost` rows next to each seam.  They
// travel over NVLink without occupying an SM and without a collective library:
//   * every slab owns a halo block in its own HBM: two staging areas per seam (double buffered by
//     exchange parity) and one 32-bit flag per seam;
//   * after the rows next to a seam have been produced (first, on a side stream, while the interior
//     of the same temporal block is still being updated), the producing GPU's copy engine pushes them
//     into the neighbour's staging area (cudaMemcpyAsync on peer memory — mapped directly inside one
//     process, through CUDA IPC between processes) and a stream memory operation
//     (cuStreamWriteValue32) publishes the exchange number in the neighbour's flag;
//   * the consumer's next block starts with one small kernel that waits for both flags (acquire
//     loads with a time-out, so a dead neighbour cannot hang the GPU) and moves the staged rows into
//     the ghost rows of the grid it is about to read.
// Double buffering makes the scheme hazard free without acknowledgements: exchange e+2 can only be
// pushed after the pusher has consumed the neighbour's exchange e+1, which the neighbour produced
// after consuming exchange e from the same staging area.
// Measured on 2 x B200 (tools/probes/peer_probe.cu, profiles/r03_peer_probe_2gpu.log): push of 6 MiB
// + flag + wait + check 56 us per round, in one process and across processes alike.
#include <cstring>

#include "hmi_internal.cuh"

namespace hmi {

namespace {

constexpr size_t kFlagStride = 128;

inline size_t stage_off(const hmi_solver *s, int side, int parity) { return (size_t)(side * 2 + parity) * s->halo_stage; }
inline size_t flag_off(const hmi_solver *s, int side) { return 4 * s->halo_stage + (size_t)side * kFlagStride; }
inline int ghost_depth(const hmi_solver *s) { return s->p.ghost_top > s->p.ghost_bottom ? s->p.ghost_top : s->p.ghost_bottom; }

struct CopyIn {
    const unsigned *flag[2];      // my flags (null = no neighbour on that side)
    const uint4 *stage[2];        // staged rows of this exchange
    uint4 *ghost[2];              // where they go: the ghost rows of the current grid
    unsigned long long n16;       // 16-byte words per side
    unsigned gen;
    unsigned *error;
};

// blockIdx.y = side.  Thread 0 of every block waits for the neighbour's push of exchange `gen`
// (about 5 s at most), then the block copies its share of the staged rows.
__global__ void halo_copyin_kernel(const CopyIn a)
{
    const int side = blockIdx.y;
    if (!a.flag[side]) return;
    __shared__ int ok;
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        unsigned cur;
        ok = 1;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(a.flag[side]) : "memory");
            if ((int)(cur - a.gen) >= 0) break;
            if (clock64() - t0 > 10000000000LL) { ok = 0; break; }
            __nanosleep(200);
        }
    }
    __syncthreads();
    if (!ok) {
        if (threadIdx.x == 0) atomicExch(a.error, 1u + (unsigned)side);
        return;
    }
    const uint4 *src = a.stage[side];
    uint4 *dst = a.ghost[side];
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n16;
         i += (unsigned long long)gridDim.x * blockDim.x)
        dst[i] = __ldcv(src + i);     // written by another GPU: do not trust a cached line
}

}  // namespace

int halo_push(hmi_solver *s, cudaStream_t st)
{
    const DriverApi *drv = driver_api();
    if (!drv) return HMI_ERR_CUDA;
    const int G = ghost_depth(s);
    const size_t bytes = (size_t)G * s->pitch * s->elem;
    s->halo_gen += 1;
    const int par = (int)(s->halo_gen & 1u);
    for (int side = 0; side < 2; ++side) {
        const int ghost = side == 0 ? s->p.ghost_top : s->p.ghost_bottom;
        if (ghost == 0) continue;
        if (!s->peer[side].peer_block) return fail(HMI_ERR_STATE, "halo side %d is not connected", side);
        // my G interior rows next to the seam -> the neighbour's staging for its ghost rows on the
        // opposite side (my upper neighbour receives its *bottom* ghost rows from me)
        const long long first = side == 0 ? s->p.ghost_top : (long long)s->p.ghost_top + s->p.rows - G;
        const char *src = s->buf[s->cur] + (size_t)first * s->pitch * s->elem;
        char *dst = s->peer[side].peer_block + stage_off(s, 1 - side, par);
        CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st));
        const CUresult r = drv->streamWriteValue32((CUstream)st, (CUdeviceptr)(s->peer[side].peer_block + flag_off(s, 1 - side)),
                                                   (cuuint32_t)s->halo_gen, CU_STREAM_WRITE_VALUE_DEFAULT);
        if (r != CUDA_SUCCESS) return fail(HMI_ERR_CUDA, "cuStreamWriteValue32 on peer memory failed (CUresult %d)", (int)r);
    }
    s->halo_state = 1;
    return HMI_OK;
}

int halo_ensure_fresh(hmi_solver *s, cudaStream_t st)
{
    if (!s->linked || s->halo_state == 0) return HMI_OK;
    if (s->side_pending && s->ev_side) {
        CUDA_TRY(cudaStreamWaitEvent(st, s->ev_side, 0));
        s->side_pending = false;
    }
    if (s->halo_state == 2) {
        const int rc = halo_push(s, st);
        if (rc != HMI_OK) return rc;
    }
    const int G = ghost_depth(s);
    CopyIn a;
    memset(&a, 0, sizeof(a));
    const int par = (int)(s->halo_gen & 1u);
    for (int side = 0; side < 2; ++side) {
        const int ghost = side == 0 ? s->p.ghost_top : s->p.ghost_bottom;
        if (ghost == 0) continue;
        a.flag[side] = reinterpret_cast<const unsigned *>(s->halo_block + flag_off(s, side));
        a.stage[side] = reinterpret_cast<const uint4 *>(s->halo_block + stage_off(s, side, par));
        const long long first = side == 0 ? 0 : (long long)s->p.ghost_top + s->p.rows;
        a.ghost[side] = reinterpret_cast<uint4 *>(s->buf[s->cur] + (size_t)first * s->pitch * s->elem);
    }
    a.n16 = (unsigned long long)G * s->pitch * s->elem / 16;
    a.gen = s->halo_gen;
    a.error = s->d_error;
    halo_copyin_kernel<<<dim3(64, 2), 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    s->launches += 1;
    s->halo_state = 0;
    return HMI_OK;
}

void halo_release(hmi_solver *s)
{
    for (int side = 0; side < 2; ++side) {
        if (s->peer[side].peer_block && s->peer[side].ipc) cudaIpcCloseMemHandle(s->peer[side].peer_block);
        s->peer[side] = HaloSide();
    }
    if (s->halo_block) cudaFree(s->halo_block);
    s->halo_block = nullptr;
    s->linked = false;
    cudaGetLastError();
}

}  // namespace hmi

using hmi::fail;

extern "C" {

int hmi_halo_setup(hmi_solver *s, hmi_halo_info *out)
{
    if (!s || !out) return fail(HMI_ERR_BAD_ARG, "null argument");
    const hmi_params &p = s->p;
    if (p.ghost_top == 0 && p.ghost_bottom == 0) return fail(HMI_ERR_BAD_ARG, "a slab without ghost rows has no halo");
    if (p.ghost_top > 0 && p.ghost_bottom > 0 && p.ghost_top != p.ghost_bottom)
        return fail(HMI_ERR_BAD_ARG, "halo links need the same ghost depth above and below (%d / %d)", p.ghost_top,
                    p.ghost_bottom);
    const int G = hmi::ghost_depth(s);
    if (p.rows < G) return fail(HMI_ERR_BAD_ARG, "the slab has fewer rows (%d) than its neighbours' ghost depth (%d)", p.rows, G);
    CUDA_TRY(cudaSetDevice(p.device));
    if (!s->halo_block) {
        s->halo_stage = (size_t)hmi::round_up((long long)G * s->pitch * s->elem, 256);
        s->halo_bytes = 4 * s->halo_stage + 2 * hmi::kFlagStride;
        // a plain cudaMalloc of its own: CUDA IPC exports whole allocations
        CUDA_TRY(cudaMalloc((void **)&s->halo_block, s->halo_bytes));
    }
    CUDA_TRY(cudaMemset(s->halo_block, 0, s->halo_bytes));
    CUDA_TRY(cudaDeviceSynchronize());
    s->halo_gen = 0;
    s->halo_state = 0;
    memset(out, 0, sizeof(*out));
    out->block = s->halo_block;
    out->block_bytes = s->halo_bytes;
    out->device = p.device;
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == sizeof(out->ipc_handle), "cudaIpcMemHandle_t is 64 bytes");
    CUDA_TRY(cudaIpcGetMemHandle(&h, s->halo_block));
    memcpy(out->ipc_handle, &h, sizeof(h));
    if (!s->own_side) {
        int lo = 0, hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&s->own_side, cudaStreamNonBlocking, hi));
    }
    return HMI_OK;
}

static int halo_connected(hmi_solver *s)
{
    const bool top_ok = s->p.ghost_top == 0 || s->peer[0].peer_block;
    const bool bot_ok = s->p.ghost_bottom == 0 || s->peer[1].peer_block;
    s->linked = top_ok && bot_ok;
    return HMI_OK;
}

int hmi_halo_connect(hmi_solver *s, int32_t side, void *peer_block, int32_t peer_device)
{
    if (!s || !peer_block) return fail(HMI_ERR_BAD_ARG, "null argument");
    if (side != 0 && side != 1) return fail(HMI_ERR_BAD_ARG, "side must be 0 (above) or 1 (below)");
    if (!s->halo_block) return fail(HMI_ERR_STATE, "hmi_halo_setup has not been called");
    if ((side == 0 ? s->p.ghost_top : s->p.ghost_bottom) == 0) return fail(HMI_ERR_BAD_ARG, "no ghost rows on side %d", side);
    CUDA_TRY(cudaSetDevice(s->p.device));
    if (peer_device != s->p.device) {
        int can = 0;
        CUDA_TRY(cudaDeviceCanAccessPeer(&can, s->p.device, peer_device));
        if (!can) return fail(HMI_ERR_UNSUPPORTED, "device %d cannot access device %d's memory", s->p.device, peer_device);
        const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return fail(HMI_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", peer_device, cudaGetErrorString(e));
        cudaGetLastError();
    }
    s->peer[side].peer_block = (char *)peer_block;
    s->peer[side].ipc = false;
    return halo_connected(s);
}

int hmi_halo_connect_ipc(hmi_solver *s, int32_t side, const unsigned char *ipc_handle)
{
    if (!s || !ipc_handle) return fail(HMI_ERR_BAD_ARG, "null argument");
    if (side != 0 && side != 1) return fail(HMI_ERR_BAD_ARG, "side must be 0 (above) or 1 (below)");
    if (!s->halo_block) return fail(HMI_ERR_STATE, "hmi_halo_setup has not been called");
    if ((side == 0 ? s->p.ghost_top : s->p.ghost_bottom) == 0) return fail(HMI_ERR_BAD_ARG, "no ghost rows on side %d", side);
    CUDA_TRY(cudaSetDevice(s->p.device));
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    void *ptr = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    s->peer[side].peer_block = (char *)ptr;
    s->peer[side].ipc = true;
    return halo_connected(s);
}

int hmi_sweeps_exchange(hmi_solver *s, int64_t n, void *main_stream, void *side_stream)
{
    if (!s) return fail(HMI_ERR_BAD_ARG, "null solver");
    if (n < 1) return fail(HMI_ERR_BAD_ARG, "hmi_sweeps_exchange needs n >= 1");
    if (!s->linked) return fail(HMI_ERR_STATE, "the slab's halo links are not connected (hmi_halo_setup / hmi_halo_connect*)");
    CUDA_TRY(cudaSetDevice(s->p.device));
    cudaStream_t ms = (cudaStream_t)main_stream;
    cudaStream_t ss = side_stream ? (cudaStream_t)side_stream : s->own_side;
    int rc = hmi::halo_ensure_fresh(s, ms);
    if (rc != HMI_OK) return rc;
    if ((rc = hmi::do_sweeps_split(s, n, ms, ss)) != HMI_OK) return rc;
    if ((rc = hmi::halo_push(s, ss)) != HMI_OK) return rc;
    CUDA_TRY(cudaEventRecord(s->ev_side, ss));
    s->side_pending = true;
    return HMI_OK;
}

}  // extern "C"
