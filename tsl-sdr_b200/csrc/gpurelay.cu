/*
 * gpurelay.cu -- the IQ relay chain of a multi-GPU receiver (C ABI: include/tslb200_gpurelay.h).
 *
 * Replaces, across GPUs, the fan-out of multifm/receiver.c:78-98 (every IQ buffer to every channel worker).  No kernels:
 * the data moves with device-to-device copies on the receiving GPU's copy engine (NVLink / NVSwitch peer access), the
 * hand-shake between the GPUs is stream-ordered 32-bit counters in page-locked host memory (cuStreamWriteValue32 /
 * cuStreamWaitValue32), shared through POSIX shared memory when every GPU has its own process.
 */
#include "../../include/tslb200_gpurelay.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

static thread_local std::string g_relay_error;

static int rerr(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_relay_error = buf;
    return code;
}

#define RCUDA(expr)                                                                                 \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return rerr(GPURELAY_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
    } while (0)

extern "C" const char *gpurelay_last_error(void) { return g_relay_error.c_str(); }

namespace {

/* one cache line per counter: "filled" = batches this rank's slot has held, "pulled" = batches this rank has copied out
 * of its parent's slot (both as seq + 1 of the latest one) */
struct SlotFlags {
    volatile uint32_t filled; uint32_t pad0[15];
    volatile uint32_t pulled; uint32_t pad1[15];
};

typedef CUresult (*StreamValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

}

struct gpurelay {
    uint32_t rank = 0, world = 1, nr_slots = 2;
    int device = 0;
    size_t slot_bytes = 0;
    std::vector<void *> slots, parent_slots;
    bool parent_is_ipc = false;
    int parent_device = -1;
    cudaStream_t cs = nullptr;                      /* copy stream: waits, the peer copy, the counter updates */
    std::vector<cudaEvent_t> ev_released;           /* per slot: local consumer done with the slot's content */
    cudaEvent_t ev_producer = nullptr;
    SlotFlags *flags = nullptr;                     /* [world][nr_slots], host view */
    CUdeviceptr flags_dev = 0;                      /* device view of the same memory */
    bool own_shm = false, registered = false;
    std::string shm_name;
    size_t shm_bytes = 0;
    StreamValue32Fn wait32 = nullptr, write32 = nullptr;

    CUdeviceptr flag_addr(uint32_t r, uint32_t s, bool pulled) const
    {
        return flags_dev + ((size_t)r * nr_slots + s) * sizeof(SlotFlags) + (pulled ? offsetof(SlotFlags, pulled) : 0);
    }
};

extern "C" size_t gpurelay_flags_bytes(uint32_t world, uint32_t nr_slots)
{
    return (size_t)world * nr_slots * sizeof(SlotFlags);
}

extern "C" int gpurelay_destroy(gpurelay_t **ph)
{
    if (!ph || !*ph) return rerr(GPURELAY_E_BADARGS, "null handle");
    gpurelay *h = *ph;
    cudaSetDevice(h->device);
    if (h->cs) { cudaStreamSynchronize(h->cs); cudaStreamDestroy(h->cs); }
    for (cudaEvent_t e : h->ev_released) if (e) cudaEventDestroy(e);
    if (h->ev_producer) cudaEventDestroy(h->ev_producer);
    if (h->parent_is_ipc) for (void *p : h->parent_slots) if (p) cudaIpcCloseMemHandle(p);
    for (void *p : h->slots) cudaFree(p);
    if (h->registered) cudaHostUnregister(h->flags);
    if (!h->shm_name.empty()) {
        if (h->flags) munmap(h->flags, h->shm_bytes);
        if (h->own_shm) shm_unlink(h->shm_name.c_str());
    }
    delete h;
    *ph = nullptr;
    return GPURELAY_OK;
}

extern "C" int gpurelay_create(gpurelay_t **ph, const gpurelay_cfg *cfg)
{
    if (!ph || !cfg) return rerr(GPURELAY_E_BADARGS, "null argument");
    *ph = nullptr;
    if (cfg->struct_size != sizeof(gpurelay_cfg)) return rerr(GPURELAY_E_BADARGS, "gpurelay_cfg size mismatch");
    if (!cfg->world || cfg->rank >= cfg->world || cfg->nr_slots < 2 || !cfg->slot_bytes || (!cfg->shm_name && !cfg->flags_host))
        return rerr(GPURELAY_E_BADARGS, "incomplete configuration");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return rerr(GPURELAY_E_NODEVICE, "no CUDA device: this library has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return rerr(GPURELAY_E_BADARGS, "bad device ordinal %d", cfg->device);
    RCUDA(cudaSetDevice(cfg->device));

    gpurelay *h = new (std::nothrow) gpurelay();
    if (!h) return rerr(GPURELAY_E_NOMEM, "out of memory");
    h->rank = cfg->rank; h->world = cfg->world; h->nr_slots = cfg->nr_slots; h->device = cfg->device;
    h->slot_bytes = (size_t)cfg->slot_bytes;
    h->shm_bytes = gpurelay_flags_bytes(h->world, h->nr_slots);

#define RFAIL(code, ...) do { rerr(code, __VA_ARGS__); gpurelay_destroy(&h); return code; } while (0)
#define RTRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) RFAIL(GPURELAY_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); } while (0)
    /* stream memory operations come from the driver; no link-time dependency on libcuda */
    cudaDriverEntryPointQueryResult qr;
    void *fn = nullptr;
    RTRY(cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qr));
    h->wait32 = (StreamValue32Fn)fn;
    RTRY(cudaGetDriverEntryPoint("cuStreamWriteValue32", &fn, cudaEnableDefault, &qr));
    h->write32 = (StreamValue32Fn)fn;
    if (!h->wait32 || !h->write32) RFAIL(GPURELAY_E_CUDA, "driver lacks stream memory operations");

    if (cfg->shm_name) {
        h->shm_name = cfg->shm_name;
        h->own_shm = h->rank == 0;
        const int fd = shm_open(cfg->shm_name, h->own_shm ? (O_CREAT | O_RDWR | O_TRUNC) : O_RDWR, 0600);
        if (fd < 0) RFAIL(GPURELAY_E_INVAL, "shm_open(%s) failed", cfg->shm_name);
        if (h->own_shm && ftruncate(fd, (off_t)h->shm_bytes)) { close(fd); RFAIL(GPURELAY_E_INVAL, "ftruncate(%s) failed", cfg->shm_name); }
        void *m = mmap(nullptr, h->shm_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (m == MAP_FAILED) RFAIL(GPURELAY_E_INVAL, "mmap(%s) failed", cfg->shm_name);
        h->flags = (SlotFlags *)m;
        if (h->own_shm) memset(m, 0, h->shm_bytes);
        RTRY(cudaHostRegister(m, h->shm_bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
        h->registered = true;
    } else {
        h->flags = (SlotFlags *)cfg->flags_host;
    }
    void *dp = nullptr;
    RTRY(cudaHostGetDevicePointer(&dp, h->flags, 0));
    h->flags_dev = (CUdeviceptr)dp;

    RTRY(cudaStreamCreateWithFlags(&h->cs, cudaStreamNonBlocking));
    RTRY(cudaEventCreateWithFlags(&h->ev_producer, cudaEventDisableTiming));
    h->slots.assign(h->nr_slots, nullptr);
    h->parent_slots.assign(h->nr_slots, nullptr);
    h->ev_released.assign(h->nr_slots, nullptr);
    for (uint32_t s = 0; s < h->nr_slots; s++) {
        RTRY(cudaMalloc(&h->slots[s], h->slot_bytes));
        RTRY(cudaEventCreateWithFlags(&h->ev_released[s], cudaEventDisableTiming));
    }
#undef RTRY
#undef RFAIL
    *ph = h;
    return GPURELAY_OK;
}

extern "C" int gpurelay_slot(gpurelay_t *h, uint32_t slot, void **d_ptr)
{
    if (!h || !d_ptr || slot >= h->nr_slots) return rerr(GPURELAY_E_BADARGS, "bad argument");
    *d_ptr = h->slots[slot];
    return GPURELAY_OK;
}

extern "C" int gpurelay_export(gpurelay_t *h, uint8_t *handles)
{
    if (!h || !handles) return rerr(GPURELAY_E_BADARGS, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == GPURELAY_IPC_HANDLE_BYTES, "IPC handle size");
    RCUDA(cudaSetDevice(h->device));
    for (uint32_t s = 0; s < h->nr_slots; s++) {
        cudaIpcMemHandle_t mh;
        RCUDA(cudaIpcGetMemHandle(&mh, h->slots[s]));
        memcpy(handles + (size_t)s * GPURELAY_IPC_HANDLE_BYTES, &mh, sizeof(mh));
    }
    return GPURELAY_OK;
}

extern "C" int gpurelay_connect_ipc(gpurelay_t *h, const uint8_t *parent_handles)
{
    if (!h || !parent_handles) return rerr(GPURELAY_E_BADARGS, "null argument");
    if (h->rank == 0) return rerr(GPURELAY_E_INVAL, "rank 0 has no parent");
    RCUDA(cudaSetDevice(h->device));
    for (uint32_t s = 0; s < h->nr_slots; s++) {
        cudaIpcMemHandle_t mh;
        memcpy(&mh, parent_handles + (size_t)s * GPURELAY_IPC_HANDLE_BYTES, sizeof(mh));
        RCUDA(cudaIpcOpenMemHandle(&h->parent_slots[s], mh, cudaIpcMemLazyEnablePeerAccess));
    }
    h->parent_is_ipc = true;
    return GPURELAY_OK;
}

extern "C" int gpurelay_connect_local(gpurelay_t *h, gpurelay_t *parent)
{
    if (!h || !parent) return rerr(GPURELAY_E_BADARGS, "null argument");
    if (h->rank == 0 || parent->rank + 1 != h->rank || parent->nr_slots != h->nr_slots)
        return rerr(GPURELAY_E_INVAL, "parent must be the previous rank of the same chain");
    RCUDA(cudaSetDevice(h->device));
    if (parent->device != h->device) {
        int can = 0;
        RCUDA(cudaDeviceCanAccessPeer(&can, h->device, parent->device));
        if (!can) return rerr(GPURELAY_E_INVAL, "device %d cannot access device %d", h->device, parent->device);
        const cudaError_t e = cudaDeviceEnablePeerAccess(parent->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return rerr(GPURELAY_E_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", parent->device, cudaGetErrorString(e));
        cudaGetLastError();
    }
    h->parent_slots = parent->slots;
    h->parent_device = parent->device;
    h->parent_is_ipc = false;
    return GPURELAY_OK;
}

#define RDRV(expr)                                                                                  \
    do {                                                                                            \
        CUresult _r = (expr);                                                                       \
        if (_r != CUDA_SUCCESS) return rerr(GPURELAY_E_CUDA, "%s:%d %s -> CUresult %d", __FILE__, __LINE__, #expr, (int)_r); \
    } while (0)

/* the slot's previous occupant (batch seq - nr_slots) is no longer needed by the local consumer nor by the next rank */
static int wait_slot_free(gpurelay *h, uint64_t seq, cudaStream_t st)
{
    if (seq < h->nr_slots) return GPURELAY_OK;
    const uint32_t s = (uint32_t)(seq % h->nr_slots);
    RCUDA(cudaStreamWaitEvent(st, h->ev_released[s], 0));
    if (h->rank + 1 < h->world)
        RDRV(h->wait32((CUstream)st, h->flag_addr(h->rank + 1, s, true), (cuuint32_t)(seq - h->nr_slots + 1), CU_STREAM_WAIT_VALUE_GEQ));
    return GPURELAY_OK;
}

extern "C" int gpurelay_acquire(gpurelay_t *h, uint64_t seq, void *producer_stream)
{
    if (!h || !producer_stream) return rerr(GPURELAY_E_BADARGS, "null argument");
    if (h->rank != 0) return rerr(GPURELAY_E_INVAL, "only the ingest rank produces batches");
    RCUDA(cudaSetDevice(h->device));
    return wait_slot_free(h, seq, (cudaStream_t)producer_stream);
}

extern "C" int gpurelay_advance(gpurelay_t *h, uint64_t seq, size_t bytes, void *producer_stream, void **ready_stream)
{
    if (!h || !ready_stream) return rerr(GPURELAY_E_BADARGS, "null argument");
    if (bytes > h->slot_bytes) return rerr(GPURELAY_E_INVAL, "%zu bytes exceed the slot capacity %zu", bytes, h->slot_bytes);
    RCUDA(cudaSetDevice(h->device));
    const uint32_t s = (uint32_t)(seq % h->nr_slots);
    const cuuint32_t stamp = (cuuint32_t)(seq + 1);
    if (h->rank == 0) {
        if (producer_stream) {
            RCUDA(cudaEventRecord(h->ev_producer, (cudaStream_t)producer_stream));
            RCUDA(cudaStreamWaitEvent(h->cs, h->ev_producer, 0));
        }
    } else {
        if (!h->parent_slots[s]) return rerr(GPURELAY_E_INVAL, "not connected to the previous rank");
        if (int rc = wait_slot_free(h, seq, h->cs)) return rc;
        RDRV(h->wait32((CUstream)h->cs, h->flag_addr(h->rank - 1, s, false), stamp, CU_STREAM_WAIT_VALUE_GEQ));
        if (bytes) {
            if (h->parent_is_ipc) RCUDA(cudaMemcpyAsync(h->slots[s], h->parent_slots[s], bytes, cudaMemcpyDefault, h->cs));
            else RCUDA(cudaMemcpyPeerAsync(h->slots[s], h->device, h->parent_slots[s], h->parent_device, bytes, h->cs));
        }
        RDRV(h->write32((CUstream)h->cs, h->flag_addr(h->rank, s, true), stamp, CU_STREAM_WRITE_VALUE_DEFAULT));
    }
    RDRV(h->write32((CUstream)h->cs, h->flag_addr(h->rank, s, false), stamp, CU_STREAM_WRITE_VALUE_DEFAULT));
    *ready_stream = (void *)h->cs;
    return GPURELAY_OK;
}

extern "C" int gpurelay_release(gpurelay_t *h, uint64_t seq, void *consumer_stream)
{
    if (!h || !consumer_stream) return rerr(GPURELAY_E_BADARGS, "null argument");
    RCUDA(cudaSetDevice(h->device));
    RCUDA(cudaEventRecord(h->ev_released[seq % h->nr_slots], (cudaStream_t)consumer_stream));
    return GPURELAY_OK;
}
