/* b200-multifm -- drop-in for `multifm <cfg.json> [<taps.json> ...]` on the file_if path
 * (multifm/multifm.c:89-174): merge the JSON files, pick the source by device.type, start, run.
 * A file replay ends cleanly at EOF (the reference aborts there). */
#include "file_if.h"

#include <stdlib.h>
#include <string.h>
#include <time.h>

int main(int argc, char *const argv[])
{
    if (argc < 2) { fprintf(stderr, "usage: %s config.json [more.json ...]\n", argv[0]); return 2; }
    char err[256];
    jnode *cfg = NULL;
    for (int i = 1; i < argc; i++) {                       /* multifm.c:105-111: config_add for every file */
        jnode *n = json_parse_file(argv[i], err, sizeof(err));
        if (!n || n->type != J_OBJ) { B200_MSG("F", "MALFORMED-CONFIG", "%s: %s", argv[i], n ? "not an object" : err); return 1; }
        if (!cfg) cfg = n; else json_merge(cfg, n);
    }
    const jnode *dev = json_get(cfg, "device");
    const char *type = NULL;
    if (!dev || json_get_string(dev, "type", &type)) { B200_MSG("F", "MISSING-DEVICE", "device.type is required"); return 1; }
    struct receiver *rx = NULL;
    if (!strcmp(type, "file")) {
        if (FAILED(file_worker_thread_new(&rx, cfg))) return 1;
    } else {
        B200_MSG("F", "UNSUPPORTED-DEVICE", "device.type '%s': live SDR sources are outside this build (file_if path only)", type);
        return 1;
    }
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    receiver_set_mute(rx, false);
    if (FAILED(receiver_start(rx))) return 1;
    receiver_drain(rx);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double secs = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    size_t outs = 0;
    for (size_t c = 0; c < rx->nr_demod_threads; c++) outs += rx->channels[c].total_nr_demod_samples;
    B200_MSG("I", "DONE", "%llu IQ samples, %zu channels, %zu PCM samples, %zu messages in %.3f s (%.1f MS/s IQ)",
             (unsigned long long)rx->total_iq_samples, rx->nr_demod_threads, outs, rx->nr_messages, secs,
             (double)rx->total_iq_samples / secs / 1e6);
    receiver_cleanup(&rx);
    json_free(cfg);
    return 0;
}
