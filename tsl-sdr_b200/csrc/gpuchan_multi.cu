/*
 * gpuchan_multi.cu -- one receiver process, several GPUs: the channels of a multifm configuration sharded over N channel
 * banks (one per device), every IQ batch delivered to all of them.  C ABI in include/tslb200_gpuchan.h (gpuchan_multi_*).
 *
 * This is receiver_sample_buf_deliver's fan-out (multifm/receiver.c:78-98: every buffer to every demod thread created at
 * :195-244) across devices.  Built purely on the public single-device ABI (gpuchan_*) and the relay chain (gpurelay_*).
 * Two ways for a batch to reach the GPUs:
 *   GPUCHAN_FANOUT_HOST   every bank copies the caller's (pinned) host batch over its own PCIe link -- N DMA engines read
 *                         the same host buffer at once; nothing crosses NVLink.  Default for host submits.
 *   GPUCHAN_FANOUT_RELAY  the batch enters device[0] only (one PCIe link, as when an SDR's DMA ring is bound to one GPU)
 *                         and travels down the NVLink relay chain on the copy engines.
 */
#include "../../include/tslb200_gpuchan.h"
#include "../../include/tslb200_gpurelay.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

struct gpuchan_multi {
    uint32_t nr = 0, fanout = GPUCHAN_FANOUT_HOST;
    uint32_t nr_channels = 0;
    size_t max_batch = 0;
    std::vector<gpuchan_t *> banks;
    std::vector<int> devices;
    std::vector<uint32_t> first, count;
    std::vector<gpurelay_t *> relays;
    std::vector<cudaStream_t> aux;          /* per device: marks the bank's progress for the relay's slot release */
    cudaStream_t s_h2d = nullptr;           /* relay mode: host -> device[0] slot */
    void *flags = nullptr;
    uint64_t seq = 0;
};

extern "C" int gpuchan_multi_destroy(gpuchan_multi_t **ph)
{
    if (!ph || !*ph) return GPUCHAN_E_BADARGS;
    gpuchan_multi *m = *ph;
    for (auto &b : m->banks) if (b) gpuchan_destroy(&b);
    for (auto &r : m->relays) if (r) gpurelay_destroy(&r);
    for (size_t i = 0; i < m->aux.size(); i++) if (m->aux[i]) { cudaSetDevice(m->devices[i]); cudaStreamDestroy(m->aux[i]); }
    if (m->s_h2d) { cudaSetDevice(m->devices[0]); cudaStreamDestroy(m->s_h2d); }
    if (m->flags) cudaFreeHost(m->flags);
    delete m;
    *ph = nullptr;
    return GPUCHAN_OK;
}

extern "C" int gpuchan_multi_create(gpuchan_multi_t **ph, const gpuchan_cfg *cfg, const int32_t *devices, uint32_t nr_devices,
                                    uint32_t fanout)
{
    if (!ph || !cfg || !devices || !nr_devices) return GPUCHAN_E_BADARGS;
    *ph = nullptr;
    if (cfg->struct_size != sizeof(gpuchan_cfg) || !cfg->nr_channels || !cfg->offset_hz) return GPUCHAN_E_BADARGS;
    if (fanout != GPUCHAN_FANOUT_HOST && fanout != GPUCHAN_FANOUT_RELAY) return GPUCHAN_E_BADARGS;
    if (nr_devices > cfg->nr_channels) nr_devices = cfg->nr_channels;          /* never an empty bank */
    gpuchan_multi *m = new (std::nothrow) gpuchan_multi();
    if (!m) return GPUCHAN_E_NOMEM;
    m->nr = nr_devices; m->fanout = fanout; m->nr_channels = cfg->nr_channels; m->max_batch = cfg->max_batch_samples;
    m->banks.assign(nr_devices, nullptr); m->relays.assign(nr_devices, nullptr); m->aux.assign(nr_devices, nullptr);
    m->devices.assign(devices, devices + nr_devices);
    m->first.resize(nr_devices); m->count.resize(nr_devices);
    /* contiguous channel ranges, sizes differing by at most one (the lower devices take the extra channels) */
    const uint32_t base = cfg->nr_channels / nr_devices, extra = cfg->nr_channels % nr_devices;
    int rc = GPUCHAN_OK;
    for (uint32_t i = 0; i < nr_devices && rc == GPUCHAN_OK; i++) {
        m->first[i] = i * base + (i < extra ? i : extra);
        m->count[i] = base + (i < extra ? 1 : 0);
        gpuchan_cfg c = *cfg;
        c.device = devices[i];
        c.nr_channels = m->count[i];
        c.offset_hz = cfg->offset_hz + m->first[i];
        c.gain = cfg->gain ? cfg->gain + m->first[i] : nullptr;
        rc = gpuchan_create(&m->banks[i], &c);
    }
    if (rc == GPUCHAN_OK && fanout == GPUCHAN_FANOUT_RELAY && nr_devices > 1) {
        const uint32_t nr_slots = 3;
        if (cudaHostAlloc(&m->flags, gpurelay_flags_bytes(nr_devices, nr_slots), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess)
            rc = GPUCHAN_E_CUDA;
        else memset(m->flags, 0, gpurelay_flags_bytes(nr_devices, nr_slots));
        for (uint32_t i = 0; i < nr_devices && rc == GPUCHAN_OK; i++) {
            gpurelay_cfg rcfg;
            memset(&rcfg, 0, sizeof(rcfg));
            rcfg.struct_size = sizeof(rcfg); rcfg.rank = i; rcfg.world = nr_devices; rcfg.nr_slots = nr_slots;
            rcfg.device = devices[i]; rcfg.slot_bytes = (uint64_t)cfg->max_batch_samples * 4; rcfg.flags_host = m->flags;
            if (gpurelay_create(&m->relays[i], &rcfg)) rc = GPUCHAN_E_CUDA;
            if (rc == GPUCHAN_OK && i > 0 && gpurelay_connect_local(m->relays[i], m->relays[i - 1])) rc = GPUCHAN_E_CUDA;
            if (rc == GPUCHAN_OK && (cudaSetDevice(devices[i]) != cudaSuccess ||
                                     cudaStreamCreateWithFlags(&m->aux[i], cudaStreamNonBlocking) != cudaSuccess)) rc = GPUCHAN_E_CUDA;
        }
        if (rc == GPUCHAN_OK && (cudaSetDevice(devices[0]) != cudaSuccess ||
                                 cudaStreamCreateWithFlags(&m->s_h2d, cudaStreamNonBlocking) != cudaSuccess)) rc = GPUCHAN_E_CUDA;
    } else if (rc == GPUCHAN_OK) {
        m->fanout = GPUCHAN_FANOUT_HOST;        /* one device: nothing to relay */
    }
    if (rc != GPUCHAN_OK) { gpuchan_multi_destroy(&m); return rc; }
    *ph = m;
    return GPUCHAN_OK;
}

extern "C" int gpuchan_multi_submit(gpuchan_multi_t *m, const int16_t *iq_host, size_t n_complex)
{
    if (!m || (!iq_host && n_complex)) return GPUCHAN_E_BADARGS;
    if (n_complex > m->max_batch) return GPUCHAN_E_INVAL;
    if (m->fanout == GPUCHAN_FANOUT_HOST) {
        /* every bank's own input stream copies the batch: N PCIe links read the same host buffer concurrently */
        for (uint32_t i = 0; i < m->nr; i++)
            if (int rc = gpuchan_submit(m->banks[i], iq_host, n_complex)) return rc;
        return GPUCHAN_OK;
    }
    const uint64_t seq = m->seq++;
    void *slot0 = nullptr;
    gpurelay_slot(m->relays[0], (uint32_t)(seq % 3), &slot0);
    if (cudaSetDevice(m->devices[0]) != cudaSuccess) return GPUCHAN_E_CUDA;
    if (gpurelay_acquire(m->relays[0], seq, m->s_h2d)) return GPUCHAN_E_CUDA;
    if (n_complex && cudaMemcpyAsync(slot0, iq_host, n_complex * 4, cudaMemcpyHostToDevice, m->s_h2d) != cudaSuccess) return GPUCHAN_E_CUDA;
    for (uint32_t i = 0; i < m->nr; i++) {
        void *ready = nullptr, *slot = nullptr;
        if (gpurelay_advance(m->relays[i], seq, n_complex * 4, i == 0 ? (void *)m->s_h2d : nullptr, &ready)) return GPUCHAN_E_CUDA;
        gpurelay_slot(m->relays[i], (uint32_t)(seq % 3), &slot);
        if (int rc = gpuchan_submit_device(m->banks[i], (const int16_t *)slot, n_complex, ready)) return rc;
        if (int rc = gpuchan_stream_wait(m->banks[i], m->aux[i])) return rc;
        if (gpurelay_release(m->relays[i], seq, m->aux[i])) return GPUCHAN_E_CUDA;
    }
    return GPUCHAN_OK;
}

extern "C" int gpuchan_multi_pending(gpuchan_multi_t *m, size_t *n_per_channel)
{
    if (!m || !n_per_channel) return GPUCHAN_E_BADARGS;
    return gpuchan_pending(m->banks[0], n_per_channel);
}

extern "C" int gpuchan_multi_collect(gpuchan_multi_t *m, int16_t *pcm_host, size_t cap_per_channel, size_t *n_per_channel)
{
    if (!m || !pcm_host || !n_per_channel) return GPUCHAN_E_BADARGS;
    *n_per_channel = 0;
    /* start every device's copy-out first (each on its own PCIe link), then wait for all of them */
    for (uint32_t i = 0; i < m->nr; i++)
        if (int rc = gpuchan_collect_begin(m->banks[i], pcm_host + (size_t)m->first[i] * cap_per_channel, cap_per_channel)) return rc;
    size_t n0 = 0;
    for (uint32_t i = 0; i < m->nr; i++) {
        size_t n = 0;
        if (int rc = gpuchan_collect_end(m->banks[i], &n)) return rc;
        if (i == 0) n0 = n;
        else if (n != n0) return GPUCHAN_E_INVAL;       /* every bank sees the same stream: cannot happen */
    }
    *n_per_channel = n0;
    return GPUCHAN_OK;
}

extern "C" int gpuchan_multi_discard(gpuchan_multi_t *m)
{
    if (!m) return GPUCHAN_E_BADARGS;
    for (uint32_t i = 0; i < m->nr; i++) gpuchan_discard(m->banks[i]);
    return GPUCHAN_OK;
}

extern "C" int gpuchan_multi_sync(gpuchan_multi_t *m)
{
    if (!m) return GPUCHAN_E_BADARGS;
    for (uint32_t i = 0; i < m->nr; i++)
        if (int rc = gpuchan_sync(m->banks[i])) return rc;
    return GPUCHAN_OK;
}

extern "C" uint32_t gpuchan_multi_devices(gpuchan_multi_t *m) { return m ? m->nr : 0; }

extern "C" int gpuchan_multi_bank(gpuchan_multi_t *m, uint32_t i, gpuchan_t **bank, uint32_t *first_channel, uint32_t *nr_channels)
{
    if (!m || i >= m->nr) return GPUCHAN_E_BADARGS;
    if (bank) *bank = m->banks[i];
    if (first_channel) *first_channel = m->first[i];
    if (nr_channels) *nr_channels = m->count[i];
    return GPUCHAN_OK;
}
