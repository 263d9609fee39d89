// Probe (not product code): do copy-engine pushes into a neighbour GPU's memory plus stream memory
// operations (cuStreamWriteValue32 / cuStreamWaitValue32) on peer memory work
//   (a) between two devices of one process (cudaDeviceEnablePeerAccess), and
//   (b) between two processes (cudaIpcOpenMemHandle)?
// and what does one push + flag + wait round trip cost?   Build: nvcc -o peer_probe peer_probe.cu -lcuda
//   ./peer_probe inproc | ipc
#include <cuda.h>
#include <cuda_runtime.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)
#define CU(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s_ = nullptr; cuGetErrorString(r_, &s_); fprintf(stderr, "%s: %s (line %d)\n", #x, s_ ? s_ : "?", __LINE__); exit(3); } } while (0)

static const size_t STAGE = 6u << 20;     // bytes pushed per exchange (24 rows x 32768 x 8 B)
static const int ITERS = 200;

__global__ void fill_kernel(unsigned *p, size_t n, unsigned v)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void check_kernel(const unsigned *p, size_t n, unsigned v, unsigned *errs)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (p[i] != v) atomicAdd(errs, 1u);
}

// kernel variants of the flag protocol: a one-thread release store into the peer's flag, and a wait
// fused into the consumer (thread 0 of every block spins with a time-out, then the block proceeds)
__global__ void signal_kernel(unsigned *flag, unsigned v)
{
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(v) : "memory");
}
__global__ void wait_check_kernel(const unsigned *flag, unsigned gen, const unsigned *p, size_t n, unsigned v, unsigned *errs)
{
    __shared__ int ok;
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        unsigned cur;
        ok = 1;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(flag) : "memory");
            if (clock64() - t0 > 4000000000LL) { ok = 0; break; }     // ~2 s
        } while ((int)(cur - gen) < 0);
    }
    __syncthreads();
    if (!ok) { if (threadIdx.x == 0) atomicAdd(errs, 1000000u); return; }
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (p[i] != v) atomicAdd(errs, 1u);
}
static int g_kernel_flags = 0;

struct Side {
    int dev;
    char *block;            // [stage0 | stage1 | flag]
    unsigned *src, *errs;
    char *peer_block;
    cudaStream_t st;
};

static void side_alloc(Side &s, int dev)
{
    s.dev = dev;
    CK(cudaSetDevice(dev));
    CK(cudaMalloc((void **)&s.block, 2 * STAGE + 256));
    CK(cudaMemset(s.block, 0, 2 * STAGE + 256));
    CK(cudaMalloc((void **)&s.src, STAGE));
    CK(cudaMalloc((void **)&s.errs, 4));
    CK(cudaMemset(s.errs, 0, 4));
    CK(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    CK(cudaDeviceSynchronize());
}

// one iteration of the protocol, enqueued only (no host sync)
static void enqueue(Side &s, int it)
{
    CK(cudaSetDevice(s.dev));
    const int par = it & 1;
    fill_kernel<<<64, 256, 0, s.st>>>(s.src, STAGE / 4, (unsigned)it);
    CK(cudaMemcpyAsync(s.peer_block + par * STAGE, s.src, STAGE, cudaMemcpyDeviceToDevice, s.st));
    if (g_kernel_flags) {
        signal_kernel<<<1, 1, 0, s.st>>>((unsigned *)(s.peer_block + 2 * STAGE), (unsigned)it);
        wait_check_kernel<<<64, 256, 0, s.st>>>((const unsigned *)(s.block + 2 * STAGE), (unsigned)it,
                                                (const unsigned *)(s.block + par * STAGE), STAGE / 4, (unsigned)it, s.errs);
        return;
    }
    CU(cuStreamWriteValue32((CUstream)s.st, (CUdeviceptr)(s.peer_block + 2 * STAGE), (cuuint32_t)it, CU_STREAM_WRITE_VALUE_DEFAULT));
    CU(cuStreamWaitValue32((CUstream)s.st, (CUdeviceptr)(s.block + 2 * STAGE), (cuuint32_t)it, CU_STREAM_WAIT_VALUE_GEQ));
    check_kernel<<<64, 256, 0, s.st>>>((const unsigned *)(s.block + par * STAGE), STAGE / 4, (unsigned)it, s.errs);
}

static unsigned side_errs(Side &s)
{
    unsigned e = 0;
    CK(cudaSetDevice(s.dev));
    CK(cudaStreamSynchronize(s.st));
    CK(cudaMemcpy(&e, s.errs, 4, cudaMemcpyDeviceToHost));
    return e;
}

static int run_inproc()
{
    int can01 = 0, can10 = 0;
    CK(cudaDeviceCanAccessPeer(&can01, 0, 1));
    CK(cudaDeviceCanAccessPeer(&can10, 1, 0));
    printf("inproc: canAccessPeer 0->1 %d, 1->0 %d\n", can01, can10);
    Side a, b;
    side_alloc(a, 0);
    side_alloc(b, 1);
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
    CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0));
    a.peer_block = b.block;
    b.peer_block = a.block;
    for (int it = 1; it <= 5; ++it) { enqueue(a, it); enqueue(b, it); }
    side_errs(a); side_errs(b);
    const auto t0 = std::chrono::steady_clock::now();
    for (int it = 6; it <= 5 + ITERS; ++it) { enqueue(a, it); enqueue(b, it); }
    const unsigned ea = side_errs(a), eb = side_errs(b);
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / ITERS;
    printf("inproc: %d iterations, errors %u / %u, %.1f us per exchange round (fill + %zu MiB push + flag + wait + check)\n",
           ITERS, ea, eb, us, STAGE >> 20);
    return (ea || eb) ? 1 : 0;
}

static int run_ipc()
{
    int sv[2];
    if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv)) { perror("socketpair"); return 2; }
    const pid_t pid = fork();
    const int me = pid == 0 ? 1 : 0;
    const int fd = sv[me];
    Side s;
    side_alloc(s, me);
    cudaIpcMemHandle_t mine, theirs;
    CK(cudaIpcGetMemHandle(&mine, s.block));
    if (write(fd, &mine, sizeof(mine)) != (ssize_t)sizeof(mine)) return 2;
    if (read(fd, &theirs, sizeof(theirs)) != (ssize_t)sizeof(theirs)) return 2;
    CK(cudaIpcOpenMemHandle((void **)&s.peer_block, theirs, cudaIpcMemLazyEnablePeerAccess));
    char ok = 1, other = 0;
    for (int it = 1; it <= 5; ++it) enqueue(s, it);
    side_errs(s);
    if (write(fd, &ok, 1) != 1 || read(fd, &other, 1) != 1) return 2;     // both warmed up
    const auto t0 = std::chrono::steady_clock::now();
    for (int it = 6; it <= 5 + ITERS; ++it) enqueue(s, it);
    const unsigned e = side_errs(s);
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / ITERS;
    printf("ipc rank %d: %d iterations, errors %u, %.1f us per exchange round\n", me, ITERS, e, us);
    fflush(stdout);
    if (write(fd, &ok, 1) != 1 || read(fd, &other, 1) != 1) return 2;     // neither unmaps while the other still pushes
    CK(cudaIpcCloseMemHandle(s.peer_block));
    if (pid == 0) _exit(e ? 1 : 0);
    int st = 0;
    waitpid(pid, &st, 0);
    return (e || st) ? 1 : 0;
}

int main(int argc, char **argv)
{
    const char *mode = argc > 1 ? argv[1] : "inproc";
    g_kernel_flags = argc > 2 && !strcmp(argv[2], "kernel");
    printf("mode %s, flags by %s\n", mode, g_kernel_flags ? "kernels (st.release.sys / spin)" : "stream memory operations");
    fflush(stdout);
    int n = 0;
    if (!strcmp(mode, "ipc")) return run_ipc();       // fork before any CUDA call
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs, found %d\n", n); return 0; }
    return run_inproc();
}
