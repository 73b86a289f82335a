/* tc_engine.cuh -- interface between gpuchan.cu (host object) and tc_engine.cu (tensor-core kernels). */
#pragma once
#include "fm_math.cuh"

#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

namespace tslb200 {

constexpr int TC_N  = 64;               /* FIR outputs (columns) per tile */
constexpr int TC_LEAD = 8;              /* columns at the head of a chunk that only provide the discriminator's previous sample */
constexpr int TC_STEP = 8;              /* columns one epilogue thread turns into PCM */
constexpr int TC_SUB = TC_N / TC_STEP;  /* derotator checkpoints per tile (one per TC_STEP columns) */
constexpr int TC_CH = 64;               /* channels per CTA (128 accumulator rows: re/im interleaved) */

struct TcPlan {
    bool ok = false;
    const char *why = "";
    int T = 0, D = 0, C = 0;
    int Kp = 0;                         /* bytes per block-row per plane: round_up(2*D, 32) */
    int Q = 0;                          /* block-rows spanned by the filter: ceil(T / D) */
    int limbs = 2;                      /* 1 when every tap entry fits in int8 */
    int R = 0;                          /* plane rows per tile: TC_N + Q - 1 */
    int G = 0;                          /* channel groups of TC_CH */
    size_t a_group_bytes = 0;           /* bytes of one group's tap image */
    size_t b_stage_bytes = 0;           /* bytes of one sample tile (both planes) */
    size_t smem_bytes = 0;
};

TcPlan tc_make_plan(int T, int D, int C, const int16_t *c_re, const int16_t *c_im, int smem_max);
/* tap image for all groups: [G][Q][limbs][Kp/16][128][16] bytes */
void tc_build_tap_image(const TcPlan &pl, const int16_t *c_re, const int16_t *c_im, std::vector<uint8_t> &img);

/* How the K outputs of one submit are cut up: `chunks` contiguous ranges of L = TC_N * n_tiles - TC_LEAD outputs
 * (one CTA per range and channel group); tile i of chunk j covers outputs j*L - TC_LEAD + TC_N*i + [0, TC_N). */
struct TcGeom {
    int chunks = 0;
    int n_tiles = 0;
    long long L = 0;
};
TcGeom tc_geometry(const TcPlan &pl, long long K, int nr_sms);
size_t tc_max_ckpt_tiles(const TcPlan &pl, long long max_K, int nr_sms);

struct TcBatch {
    InWindow in;
    const uint8_t *tap_img;
    const int *incr, *ckpt, *last_in;
    int *last_out;
    const float2 *atan_tab;
    short *pcm;
    int *iq_out;
    long long pitch;
    unsigned long long K;
    TcGeom geom;
    AtanParams atan;
    long long *dbg = nullptr;
    int dbg_flags = 0;
};

cudaError_t tc_launch_fir_fm(const TcPlan &pl, const TcBatch &b, cudaStream_t st);

} // namespace tslb200
