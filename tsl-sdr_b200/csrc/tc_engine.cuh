/* tc_engine.cuh -- interface between gpuchan.cu (host object) and tc_engine.cu (tensor-core kernels). */
#pragma once
#include "fm_math.cuh"

#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

namespace tslb200 {

constexpr int TC_OUT  = 64;             /* PCM outputs (columns that are turned into PCM) per tile */
constexpr int TC_LEAD = 16;             /* columns computed ahead of them; only the last one is used (the discriminator's
                                           previous sample), 16 because M = 128 MMAs need N % 16 == 0 */
constexpr int TC_N    = TC_OUT + TC_LEAD;   /* FIR outputs (MMA N) per tile */
constexpr int TC_STEP = 8;              /* consecutive outputs one epilogue thread turns into PCM (one 16-byte store) */
constexpr int TC_SUB  = TC_OUT / TC_STEP;   /* derotator checkpoints per tile (one per TC_STEP outputs) */
constexpr int TC_CH   = 64;             /* channels per CTA (128 accumulator rows) */
constexpr int TC_ACC_STRIDE = TC_N;     /* TMEM columns between the limb accumulators of one stage */
constexpr int TC_PROG_MAX = 160;        /* MMA instructions per tile the kernel parameter block can describe */

enum { TC_MODE_SUM = 0, TC_MODE_RADIX = 1 };

/* One tcgen05.mma of the per-tile program, pre-digested on the host so that the issuing warp needs one 16-byte
 * constant load and two adds per instruction (the uniform datapath that feeds UTCIMMA is slow):
 *   a_lo  = low word of the A (tap image) shared-memory descriptor without the image base: chunk offset in 16-byte
 *           units | LBO << 16;  the kernel adds smem_addr(image) >> 4
 *   b_lo  = same for the B (sample stage) operand: (plane, K chunk, row shift) offset | LBO << 16;  + stage base
 *   d_acc = TMEM column offset of the accumulator inside a stage | accumulate << 31 (0 for the first MMA into it)
 *   idesc = the kind::i8 instruction descriptor (operand signedness, M, N) */
struct TcMma {
    uint32_t a_lo, b_lo, d_acc, idesc;
};

struct TcPlan {
    bool ok = false;
    const char *why = "";
    int T = 0, D = 0, C = 0;
    int Kp = 0;                         /* bytes per block-row per plane: round_up(2*D, 32) */
    int Q = 0;                          /* block-rows spanned by the filter: ceil(T / D) */
    int R = 0;                          /* plane rows per tile: TC_N + Q - 1 */
    int G = 0;                          /* channel groups of TC_CH */
    int gpc = 1;                        /* channel groups per CTA (2: a transformed sample tile feeds two groups' MMAs) */
    int mode = TC_MODE_RADIX;           /* how int16 taps are made of int8 operands */
    int accs = 3;                       /* limb accumulators per TMEM stage (SUM: 2, RADIX: 3) */
    int nb_stages = 2;                  /* sample stages in shared memory */
    int nt_stages = 2;                  /* accumulator stages in TMEM */
    int a_chunks = 0;                   /* 4 KB chunks in one group's tap image */
    size_t a_group_bytes = 0;           /* bytes of one group's tap image */
    size_t b_stage_bytes = 0;           /* bytes of one sample tile (both planes) */
    size_t smem_bytes = 0;
    int atan_copies = 1;                /* interleaved copies of the arctangent table in shared memory (16 or 1) */
    std::vector<TcMma> prog;            /* the MMAs of one tile: [0, prog_split) issued by MMA warp 0, the rest by warp 1 */
    bool prog_regular = false;          /* (accumulator, descriptor) alternate even/odd within each warp's part: fast issue path */
    int prog_split = 0;                 /* the two warps own disjoint accumulators, so their order does not matter */
    /* tap image construction: for image chunk i, which (q, kk) it covers and which limb/term it holds */
    struct Chunk { int q, kk, term; };
    std::vector<Chunk> chunks;
};

TcPlan tc_make_plan(int T, int D, int C, const int16_t *c_re, const int16_t *c_im, int smem_max);
/* tap image for all groups: [G][a_chunks][2 slabs][128 rows][16] bytes */
void tc_build_tap_image(const TcPlan &pl, const int16_t *c_re, const int16_t *c_im, std::vector<uint8_t> &img);

/* How the K outputs of one submit are cut up: tile t covers outputs TC_OUT*t + [0, TC_OUT) (plus TC_LEAD lead-in
 * columns before them); CTA j of a channel group walks tiles [j * n_tiles, (j + 1) * n_tiles). */
struct TcGeom {
    int chunks = 0;                     /* CTAs per channel group */
    int n_tiles = 0;                    /* tiles per CTA */
    int total_tiles = 0;
};
TcGeom tc_geometry(const TcPlan &pl, long long K, int nr_sms);
size_t tc_max_ckpt_tiles(const TcPlan &pl, long long max_K, int nr_sms);

struct TcBatch {
    InWindow in;
    const uint8_t *tap_img;
    const int *incr, *ckpt, *last_in;   /* ckpt == nullptr: derotator phases come straight from the cycle table */
    const uint32_t *mu = nullptr, *lambda = nullptr;    /* per channel: transient length and period of the derotator */
    const int *cyc = nullptr;           /* [C][cyc_pitch] one period of every channel's derotator sequence */
    int cyc_pitch = 0;
    unsigned long long k_base = 0;      /* stream index of this submit's output 0 */
    int *carry_out = nullptr;           /* where to keep the input samples the next submit still needs ... */
    long long carry_from = 0;           /* ... [carry_from, carry_from + carry_keep) of this submit's window */
    int carry_keep = 0;
    int *last_out;
    const float2 *atan_tab;
    short *pcm;
    int *iq_out;
    long long pitch;
    unsigned long long K;
    TcGeom geom;
    AtanParams atan;
};

cudaError_t tc_launch_fir_fm(const TcPlan &pl, const TcBatch &b, cudaStream_t st);

} // namespace tslb200
