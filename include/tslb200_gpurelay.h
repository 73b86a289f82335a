/*
 * tslb200_gpurelay.h -- C ABI of the IQ relay: the one data-path exchange of a multi-GPU receiver.
 *
 * The reference hands every IQ sample_buf to every channel worker (multifm/receiver.c:78-98, refcount = number of
 * demod threads).  With the channels sharded over several GPUs the same fan-out is "every GPU needs every batch":
 * the batch enters one GPU (the ingest GPU, relay rank 0) and travels down a CHAIN 0 -> 1 -> ... -> N-1 over
 * NVLink.  Every GPU sends once and receives once per batch, all hops of consecutive batches run at the same time, so
 * the steady-state cost per batch is one hop (bytes / NVLink bandwidth) whatever N is; a broadcast from one root would
 * cost N-1 hops of the root's egress.  The copies are plain device-to-device copies on the receiving GPU's copy
 * engine: no SM is used, the persistent channel-bank kernels run undisturbed next to them (NCCL's broadcast kernels
 * slowed them by 50 %, profiles/r01_bench_n2.json).
 *
 * Synchronisation between the GPUs is stream-ordered: per (rank, slot) a 32-bit "filled" and "pulled" counter in
 * page-locked host memory shared by all ranks, written with cuStreamWriteValue32 when a copy has completed and waited
 * for with cuStreamWaitValue32 -- no host thread ever blocks, and it works across processes (one process per GPU under
 * torchrun: the counters live in a POSIX shared-memory segment, the slot buffers are opened through CUDA IPC) exactly as
 * inside one process (gpuchan_multi_*, tslb200_gpuchan.h).
 *
 * Batch number seq (0, 1, 2, ...; every rank must be driven with the same sequence) lives in slot seq % nr_slots of every
 * rank.  A slot is overwritten only after the local consumer released its previous content (gpurelay_release) and the next
 * rank of the chain pulled it.
 */
#ifndef TSLB200_GPURELAY_H
#define TSLB200_GPURELAY_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPURELAY_OK            0
#define GPURELAY_E_NOMEM     (-1)
#define GPURELAY_E_BADARGS   (-2)
#define GPURELAY_E_INVAL     (-5)
#define GPURELAY_E_CUDA      (-64)
#define GPURELAY_E_NODEVICE  (-65)

#define GPURELAY_IPC_HANDLE_BYTES 64

typedef struct gpurelay gpurelay_t;

typedef struct gpurelay_cfg {
    uint32_t struct_size;
    uint32_t rank, world;        /* position in the chain, chain length */
    uint32_t nr_slots;           /* batches in flight along the chain, >= 2 */
    int32_t  device;             /* CUDA ordinal of this rank's GPU */
    uint32_t reserved;
    uint64_t slot_bytes;         /* capacity of one slot */
    const char *shm_name;        /* POSIX shared-memory name of the counters ("/name"); rank 0 creates it, the others open it
                                    (after rank 0's gpurelay_create returned).  NULL: counters = flags_host */
    void *flags_host;            /* one process driving all GPUs: gpurelay_flags_bytes(world, nr_slots) bytes of page-locked,
                                    zeroed host memory shared by all ranks' relay objects */
} gpurelay_cfg;

size_t gpurelay_flags_bytes(uint32_t world, uint32_t nr_slots);

int gpurelay_create(gpurelay_t **ph, const gpurelay_cfg *cfg);
int gpurelay_destroy(gpurelay_t **ph);

/* Device address of a slot of this rank. */
int gpurelay_slot(gpurelay_t *h, uint32_t slot, void **d_ptr);

/* Connect to the previous rank of the chain (not for rank 0).
 * Across processes: rank r-1 exports its slots (nr_slots x GPURELAY_IPC_HANDLE_BYTES bytes), the bytes travel by any
 * means (torch.distributed in bench.py), rank r opens them.  Inside one process: pass the parent's object. */
int gpurelay_export(gpurelay_t *h, uint8_t *handles);
int gpurelay_connect_ipc(gpurelay_t *h, const uint8_t *parent_handles);
int gpurelay_connect_local(gpurelay_t *h, gpurelay_t *parent);

/* Rank 0 only: make producer_stream wait until slot seq % nr_slots may be overwritten with batch seq. */
int gpurelay_acquire(gpurelay_t *h, uint64_t seq, void *producer_stream);

/* Bring batch seq (`bytes` bytes) into this rank's slot and publish it to the next rank.
 *   rank 0 : the batch was produced into the slot by work on producer_stream (NULL: it is already there);
 *   rank r : waits until rank r-1 holds batch seq, copies it over NVLink on this GPU's copy engine.
 * *ready_stream receives the stream (cudaStream_t) after whose current position the slot holds the batch on this GPU:
 * hand it to gpuchan_submit_device. */
int gpurelay_advance(gpurelay_t *h, uint64_t seq, size_t bytes, void *producer_stream, void **ready_stream);

/* The local consumer of batch seq has been enqueued on consumer_stream: its current position is the point after which this
 * GPU no longer reads the slot. */
int gpurelay_release(gpurelay_t *h, uint64_t seq, void *consumer_stream);

const char *gpurelay_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
