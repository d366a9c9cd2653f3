// Tensor-core shading path (north-star kernel 2): SpatialEncoder + GeoVisFusion + MLPUNetFusion + TexVisFusion +
// IBRRenderingHead + eval_func for tiles of 128 samples, bf16 operands, fp32 accumulation, on tcgen05 / TMEM.
//   SpatialEncoder.forward (rel_z_decay)   src/spatial.py:59-117
//   GeoVisFusion.forward                   src/networks.py:75-106
//   MLPUNetFusion.forward                  src/utils.py:633-649 (MLPUNet :822-852, PoolModule :744-779, pool_ops :854-880)
//   ibr_compress_gfeat + TexVisFusion      src/model.py:921, src/networks.py:281-293
//   IBRRenderingHead.forward               src/model.py:1600-1636
//   eval_func                              src/model.py:1140-1160
//
// One CTA per SM works on TWO 128-sample tiles at a time (persistent over tile pairs): each tile has its own group of
// 8 warps (two threads per row: warps w and w+4 of the group share the TMEM lane quarter w and split the columns), its
// own five 16 KB operand slots (128 rows x 64 bf16, K-major, 128B swizzle), its own accumulator / pooling TMEM columns
// ([0,128) / [128,256) of its 256-column half), its own MMA issuer warp and its own barriers, and runs the layer
// sequence independently, so one tile's epilogue can overlap the other's MMAs.  Both tiles consume the SAME weight
// stream: one producer lane brings the weight images in once per tile pair through a 4 x 16 KB ring (cp.async.bulk +
// mbarriers; a ring slot is released when the MMAs of both tiles that read it have completed).  Activations never leave
// the SM.  Biases, the small fp32 layers and the camera-space keypoints travel as a kernel parameter (constant bank).
// The layer sequence is a table of "steps" (a set of MMAs whose results are consumed by one epilogue), known at compile
// time (kProg) and rebuilt on the host together with the weight images, so that the packer, the producer and the issuer
// cannot disagree.  Code size is a first-order cost of this kernel (it streams its instructions from L2): see the notes
// at TC_ROLL_VIEWS / TC_ROLL_MLP / tc_issuer_warp and DESIGN.md section 5.
#pragma once
#include <algorithm>
#include "common.cuh"
#include "gather_tc.cuh"
#include "tc_prims.cuh"

#define TC_NACT 5
// Weight ring depth.  With the tables passed as a kernel parameter (constant bank, TC_TAB_PARAM) the shared memory they
// used to take is enough for a fourth 16 KB ring slot: a whole 3-chunk step plus the first chunk of the next one.
#ifndef TC_TAB_PARAM
#define TC_TAB_PARAM 1
#endif
#ifndef TC_NRING
#if TC_TAB_PARAM && !defined(VANERF_TC_TRACE)
#define TC_NRING 4
#else
#define TC_NRING 3
#endif
#endif
#define TC_TILES 2                       // tiles in flight per CTA
#define TC_EPI_THREADS 256               // threads of one tile group
#define TC_THREADS (TC_TILES * TC_EPI_THREADS + 128)   // + one warpgroup: weight producer warp, one MMA issuer warp per tile, one parked warp
#ifndef TC_REGS_EPI
#define TC_REGS_EPI 112                  // registers of a tile thread after setmaxnreg (kernel is compiled for 96)
#endif
#define TC_REGS_PROD 32                  // registers of the producer / issuer warpgroup (gives 64 x 128 back to the pool)
#define TC_NREADY 4                      // operand-ready barriers per tile: a warp runs at most 3 steps ahead of the slowest (PE ring)
#define TC_TMEM_COLS 512
#define TC_TMEM_TILE 256                 // TMEM columns per tile
#define TC_SREG 128                      // TMEM columns of the pooling sums S1|S2, later the per-view f (40 each)
#define TC_SRC 112                       // TMEM columns of the per-view source colours (4 per view)
#define TC_MAXV 3
// Split-precision ("fp32") variant of the kernel, template parameter SPLIT: every operand is kept as bf16 hi + bf16 lo
// (x = hi + lo to 16 mantissa bits) and every K step issues three MMAs (hi*hi + lo*hi + hi*lo; the dropped lo*lo term is
// 2^-18 relative), accumulating in fp32 in TMEM.  One 128-sample tile per CTA: the second tile's operand slots hold the lo
// images, the weight ring carries (hi, lo) chunk pairs.  Epilogue functions run in fp32 (no packed-bf16 shortcuts).
#define TC_LO_OFF (TC_NACT * TC_SLOT)    // byte offset of the lo image of an operand slot (= the other tile's slots)
#define TC_AUX_BYTES_SPLIT 96            // side record of the split path: the 8 "extras" travel as fp32
#define TC_OFF_RING (TC_TILES * TC_NACT * TC_SLOT)
#define TC_OFF_TAB (TC_OFF_RING + TC_NRING * TC_SLOT)
#if TC_TAB_PARAM
#define TC_TAB_BYTES 0                   // tables live in the kernel parameter (constant bank)
#else
#define TC_TAB_BYTES 6656                // >= sizeof(TcTables), multiple of 16
#endif
#define TC_OFF_CTRL (TC_OFF_TAB + TC_TAB_BYTES)
#define TC_OFF_TRACE (TC_OFF_CTRL + 512)
#ifdef VANERF_TC_TRACE
#define TC_TRACE_N 1024                   // cycle-trace entries (tag << 48 | clock), developer aid: 512 per traced thread
#else
#define TC_TRACE_N 0
#endif
#define TC_SMEM_BYTES (TC_OFF_TRACE + TC_TRACE_N * 8)

#ifndef TC_ROLL_MLP
#define TC_ROLL_MLP 1                    // the three 128-wide Softplus epilogues of MLPUNet layers1 share one loop body
#endif
// Per-view loops of the rendering head: rolled (one copy of each epilogue body, iterations 2.. hit the instruction
// cache) or unrolled.  Register arrays indexed by the view go through tc_sel3 / tc_set3 so that rolling is legal.
#ifndef TC_ROLL_VIEWS
#define TC_ROLL_VIEWS 1
#endif
#if TC_ROLL_VIEWS
#define TC_VLOOP _Pragma("unroll 1")
#else
#define TC_VLOOP _Pragma("unroll")
#endif
// Measured and dropped (round 2): starting tile group 1 half a step after group 0 (the tiles do not contend for a saturated
// resource, so their phase does not matter: 88.8 vs 90.6 ms); requesting the next view's operand images two layers early
// (h2 moved to slots 4 | 0 to free slots 1-3: 93.7 vs 90.7 ms).
#ifndef TC_F32X2
#define TC_F32X2 1                       // packed fp32 pairs (FADD2 / FMUL2 / FFMA2) in the epilogues: -4.5 % kernel time
#endif
#ifndef TC_PREWAIT
#define TC_PREWAIT 1                     // issuer waits for a step's weight chunks before it waits for the step's operands: -9 %
#endif
// Measured and dropped at the end of round 2 (the variants are in the history of this file): operand images
// requested one step before they are needed (80.1 .. 81.4 vs 79.2 ms per view); CTAs started out of phase (no effect); the four
// keypoints of a positional-encoding block as independent instruction streams (+2 %: the blocks are paced by the issuer, not by their
// generation); MLP layer 0 in two publishes with per-slot ring commits instead of seven (75.5 vs 74.5); out_layer's second Linear in
// registers on one thread per row (77.2 vs 74.5); the 64 -> 2 and 32 -> 1 Linears as partial dot products swapped between the two
// threads of a row through TMEM at a 64-thread named barrier (74.8 / 74.6 vs 74.3: a pair exchange costs what the MMA round trip
// costs); 8 .. 1 000 padding instructions in front of the body (+-0.7 %: code placement is the noise floor of all of these).
#ifndef TC_ABLATE
#define TC_ABLATE 0                      // developer timing experiments (results are wrong when non-zero): 1 softplus -> relu,
#endif                                   // 2 PE without MUFU, 4 one K step per MMA op, 8 ELU / sigmoid -> identity, 16 no gating,
                                         // 32 no proxy fence, 64 no waits for weight chunks, 128 no operand-image loads
#ifndef TC_MERGE_G4
#define TC_MERGE_G4 1                    // GeoVisFusion's last Linear (no bias, no activation) folded into the layers that consume it
#endif                                   // (MLP layer 0's out64 columns, MLP layer 2's out8 columns): no step G4, one round trip less per view
#define TC_M0_SLOT (TC_MERGE_G4 ? 4 : 0) // operand slot MLP layer 0 reads its 64 GeoVisFusion columns from (G3's output / G4's output)
#define TC_PE_SLOT2 (TC_MERGE_G4 ? 0 : 4)   // third slot of the positional-encoding ring (slots 1, 2 and this one)
#ifndef TC_I9_REGS
#define TC_I9_REGS 1                     // out_layer's last Linear (8 -> 1) in fp32 registers inside the epilogue of step I8: one round trip less per tile
#endif
enum TcStepId {
    ST_G1 = 0, ST_G2, ST_G3, ST_G4, ST_M0, ST_P0, ST_P1, ST_P2, ST_P3, ST_P4, ST_P5, ST_M1, ST_M2, ST_M3,
    ST_Q1, ST_Q2, ST_Q3, ST_T1, ST_T2, ST_T3, ST_T4, ST_I1, ST_I2, ST_I3, ST_I4, ST_I5, ST_I6, ST_I7, ST_I8, ST_I9,
    ST_COUNT
};

// steps that exist in the tables (the packer keeps their layers) but are evaluated in registers, never issued as MMAs
constexpr bool tc_step_in_regs(int st);
struct TcOp {
    uint32_t a_off;        // byte offset of the first K step inside the activation area (slot * 16 KB + (col0/16) * 32)
    uint32_t b_off;        // byte offset of the weight block inside its ring slot
    uint32_t idesc;
    uint16_t d_col;        // TMEM accumulator column
    uint8_t nk, accum, chunk_rel, last_in_chunk;
};
struct TcStep { uint16_t op0, nops, chunk0, nchunks; };
constexpr bool tc_step_in_regs(int st) { return st == ST_G2 || (TC_MERGE_G4 && st == ST_G4) || (TC_I9_REGS && st == ST_I9); }
struct TcChunk { uint32_t src_off, bytes; };
#define TC_MAX_OPS 96
#define TC_MAX_CHUNKS 64
#define TC_MAX_BIAS 1024
struct TcTables {                        // global memory (context-owned); copied to shared memory once per CTA
    alignas(16) float bias[TC_MAX_BIAS]; // read as float4 (offsets are multiples of 16 floats)
    alignas(16) float kpt4[TC_MAXV * NKPT * 4];   // keypoints in each source camera frame (per frame), xyz + pad
    alignas(16) float at2[2 * 3 * 12];   // GeoVisFusion attention layer 2 (3 x 10, rows padded to 12), scale 64 then 8: fp32, in registers
    alignas(16) float out2w[8];          // IBRRenderingHead out_layer, last Linear (8 -> 1): fp32, in registers (TC_I9_REGS)
    uint16_t bias_off[L_COUNT + 1];      // offset of each biased layer's bias in `bias`
    float ani_al_abs;
};
// The static MMA program (depends on the layer shapes only).  It lives in __constant__ memory: the issuing warp reads
// it with uniform loads, so operand descriptors are formed in uniform registers and tcgen05.mma issues back to back.
struct TcProg {
    TcStep steps[ST_COUNT];
    TcOp ops[TC_MAX_OPS];
    TcChunk chunks[TC_MAX_CHUNKS];
    uint16_t cc_off[ST_COUNT];           // weight chunks that precede the step inside one iteration of its group
    uint16_t cc_gm, cc_q, cc_t, cc_i;    // chunks per iteration of: geometry+MLP (per view, ST_G2 excluded), density head,
                                         // texture fusion (per view), rendering head (per view)
};

// ---- compile-time copy of the program --------------------------------------------------------------------------
// The layer shapes are fixed (configs/vanerf.json, checked by vanerf_load_weights), so the step / op / chunk tables are
// known at compile time.  kProg is what the MMA issuer warps execute: every descriptor offset, accumulator column and
// instruction descriptor becomes an immediate, and issuing a step costs no table loads.  The host builds the same
// tables again from the packing script (tc_build, which also produces the weight images) and tc_program_matches()
// refuses to run if the two ever disagree.
struct TcSpecC { int step, layer, slot, col0, d_col, accum, ncols, share; };   // share = k > 0: reuse the weight block of the op k specs earlier
constexpr int kLayerOut[L_COUNT] = {10, 3, 64, 64, 10, 3, 8, 8, 128, 128, 120, 64, 64, 64, 2, 24, 96, 6, 96, 40,
                                    16, 40, 64, 32, 32, 33, 32, 1, 16, 8, 1};
constexpr TcSpecC kSpecs[] = {
    {ST_G1, L_GEO_AT0, 0, 0, 0, 0, 64, 0}, {ST_G1, L_GEO_AT0, 1, 0, 0, 1, 64, 0}, {ST_G1, L_GEO_AT0, 2, 0, 0, 1, 64, 0}, {ST_G1, L_GEO_AT0, 3, 0, 0, 1, 16, 0},
    {ST_G1, L_GEO8_AT0, 3, 16, 16, 0, 32, 0},
    {ST_G2, L_GEO_AT1, 3, 48, 0, 0, 16, 0}, {ST_G2, L_GEO8_AT1, 4, 0, 16, 0, 16, 0},
    {ST_G3, L_GEO_F0, 0, 0, 0, 0, 64, 0}, {ST_G3, L_GEO_F0, 1, 0, 0, 1, 64, 0}, {ST_G3, L_GEO_F0, 2, 0, 0, 1, 64, 0}, {ST_G3, L_GEO_F0, 3, 0, 0, 1, 16, 0},
    {ST_G3, L_GEO8_F0, 3, 16, 64, 0, 32, 0},
    {ST_G4, L_GEO_F1, 4, 0, 0, 0, 64, 0}, {ST_G4, L_GEO8_F1, 3, 48, 64, 0, 16, 0},
    {ST_M0, L_MLP0, TC_M0_SLOT, 0, 0, 0, 64, 0},
    {ST_P0, L_MLP0, 1, 0, 0, 1, 64, 0}, {ST_P1, L_MLP0, 2, 0, 0, 1, 64, 0}, {ST_P2, L_MLP0, TC_PE_SLOT2, 0, 0, 1, 64, 0},
    {ST_P3, L_MLP0, 1, 0, 0, 1, 64, 0}, {ST_P4, L_MLP0, 2, 0, 0, 1, 64, 0}, {ST_P5, L_MLP0, TC_PE_SLOT2, 0, 0, 1, 16, 0},
    {ST_M1, L_MLP1, 1, 0, 0, 0, 64, 0}, {ST_M1, L_MLP1, 2, 0, 0, 1, 64, 0},
    {ST_M2, L_MLP2, 4, 0, 0, 0, 64, 0}, {ST_M2, L_MLP2, 0, 0, 0, 1, 64, 0}, {ST_M2, L_MLP2, 3, 48, 0, 1, 16, 0},
    {ST_M3, L_MLP3, 1, 0, 0, 0, 64, 0}, {ST_M3, L_MLP3, 2, 0, 0, 1, 64, 0},
    {ST_Q1, L_POST0, 1, 0, 0, 0, 64, 0}, {ST_Q1, L_POST0, 2, 0, 0, 1, 64, 0}, {ST_Q1, L_COMPRESS, 1, 0, 64, 0, 64, 0}, {ST_Q1, L_COMPRESS, 2, 0, 64, 1, 64, 0},
    {ST_Q2, L_POST1, 4, 0, 0, 0, 64, 0},
    {ST_Q3, L_POST2, 0, 0, 0, 0, 64, 0},
    {ST_T1, L_TEX_AT0, 1, 0, 0, 0, 64, 0}, {ST_T1, L_TEX_AT0, 2, 0, 0, 1, 32, 0}, {ST_T1, L_RAY0, 3, 0, 96, 0, 16, 0},
    {ST_T2, L_TEX_AT1, 4, 0, 0, 0, 64, 0}, {ST_T2, L_TEX_AT1, 0, 0, 0, 1, 32, 0}, {ST_T2, L_RAY1, 3, 16, 16, 0, 16, 0},
    {ST_T3, L_TEX_F0, 1, 0, 0, 0, 64, 0}, {ST_T3, L_TEX_F0, 2, 0, 0, 1, 32, 0},
    {ST_T4, L_TEX_F1, 4, 0, 0, 0, 64, 0}, {ST_T4, L_TEX_F1, 0, 0, 0, 1, 32, 0},
    // IBR head, the three views in one step: view v works in slot 2 + v and accumulator columns stride * v; views 1, 2
    // reuse the weight blocks of view 0 (share = distance back to the op that owns the block)
    {ST_I1, L_BASE0, 1, 0, 0, 0, 64, 0}, {ST_I1, L_BASE0, 2, 0, 0, 1, 64, 0},
    {ST_I1, L_BASE0, 1, 0, 64, 0, 64, 2}, {ST_I1, L_BASE0, 3, 0, 64, 1, 64, 2},
    {ST_I1, L_BASE0, 1, 0, 128, 0, 64, 4}, {ST_I1, L_BASE0, 4, 0, 128, 1, 64, 4},
    {ST_I2, L_BASE1, 2, 0, 0, 0, 64, 0}, {ST_I2, L_BASE1, 3, 0, 32, 0, 64, 1}, {ST_I2, L_BASE1, 4, 0, 64, 0, 64, 2},
    {ST_I3, L_VIS1_0, 2, 0, 0, 0, 32, 0}, {ST_I3, L_VIS1_0, 3, 0, 48, 0, 32, 1}, {ST_I3, L_VIS1_0, 4, 0, 96, 0, 32, 2},
    {ST_I4, L_VIS1_1, 2, 32, 0, 0, 32, 0}, {ST_I4, L_VIS1_1, 3, 32, 48, 0, 32, 1}, {ST_I4, L_VIS1_1, 4, 32, 96, 0, 32, 2},
    {ST_I5, L_VIS2_0, 2, 0, 0, 0, 32, 0}, {ST_I5, L_VIS2_0, 3, 0, 48, 0, 32, 1}, {ST_I5, L_VIS2_0, 4, 0, 96, 0, 32, 2},
    {ST_I6, L_VIS2_1, 2, 32, 0, 0, 32, 0}, {ST_I6, L_VIS2_1, 3, 32, 48, 0, 32, 1}, {ST_I6, L_VIS2_1, 4, 32, 96, 0, 32, 2},
    {ST_I7, L_OUT0, 2, 0, 0, 0, 48, 0}, {ST_I7, L_OUT0, 3, 0, 16, 0, 48, 1}, {ST_I7, L_OUT0, 4, 0, 32, 0, 48, 2},
    {ST_I8, L_OUT1, 2, 48, 0, 0, 16, 0}, {ST_I8, L_OUT1, 3, 48, 16, 0, 16, 1}, {ST_I8, L_OUT1, 4, 48, 32, 0, 16, 2},
    {ST_I9, L_OUT2, 2, 0, 0, 0, 16, 0}, {ST_I9, L_OUT2, 3, 0, 16, 0, 16, 1}, {ST_I9, L_OUT2, 4, 0, 32, 0, 16, 2},
};
constexpr int kNumSpecs = (int)(sizeof(kSpecs) / sizeof(kSpecs[0]));

constexpr TcProg tc_make_prog() {
    TcProg P{};
    int n_ops = 0, n_chunks = 0;
    uint32_t blob_bytes = 0;
    for (int s = 0; s < ST_COUNT; ++s) {
        P.steps[s].op0 = (uint16_t)n_ops;
        P.steps[s].chunk0 = (uint16_t)n_chunks;
        uint32_t cur_bytes = 0;
        int rel = -1;
        for (int q = 0; q < kNumSpecs; ++q) {
            if (kSpecs[q].step != s) continue;
            const int n_pad = (kLayerOut[kSpecs[q].layer] + 15) & ~15;
            const uint32_t bytes = (uint32_t)n_pad * 128;
            const int share = kSpecs[q].share;               // > 0: the weight block of the op `share` ops earlier
            if (!share && (rel < 0 || cur_bytes + bytes > TC_SLOT)) {
                if (rel >= 0) P.ops[n_ops - 1].last_in_chunk = 1;
                ++rel;
                P.chunks[n_chunks].src_off = blob_bytes;
                P.chunks[n_chunks].bytes = 0;
                ++n_chunks;
                cur_bytes = 0;
            }
            P.ops[n_ops].a_off = (uint32_t)(kSpecs[q].slot * TC_SLOT + (kSpecs[q].col0 / 16) * 32);
            P.ops[n_ops].b_off = share ? P.ops[n_ops - share].b_off : cur_bytes;
            P.ops[n_ops].idesc = tc::umma_idesc_bf16(128, n_pad);
            P.ops[n_ops].d_col = (uint16_t)kSpecs[q].d_col;
            P.ops[n_ops].nk = (uint8_t)(kSpecs[q].ncols / 16);
            P.ops[n_ops].accum = (uint8_t)kSpecs[q].accum;
            P.ops[n_ops].chunk_rel = share ? P.ops[n_ops - share].chunk_rel : (uint8_t)rel;
            P.ops[n_ops].last_in_chunk = 0;
            ++n_ops;
            if (!share) {
                cur_bytes += bytes;
                blob_bytes += bytes;
                P.chunks[n_chunks - 1].bytes = cur_bytes;
            }
        }
        if (n_ops > P.steps[s].op0) P.ops[n_ops - 1].last_in_chunk = 1;
        P.steps[s].nops = (uint16_t)(n_ops - P.steps[s].op0);
        P.steps[s].nchunks = (uint16_t)(n_chunks - P.steps[s].chunk0);
    }
    const int grp_first[4] = {ST_G1, ST_Q1, ST_T1, ST_I1}, grp_last[4] = {ST_M3, ST_Q3, ST_T4, ST_I9};
    for (int g = 0; g < 4; ++g) {
        int acc = 0;
        for (int st = grp_first[g]; st <= grp_last[g]; ++st) {
            P.cc_off[st] = (uint16_t)acc;
            if (!tc_step_in_regs(st)) acc += P.steps[st].nchunks;
        }
        if (g == 0) P.cc_gm = (uint16_t)acc;
        else if (g == 1) P.cc_q = (uint16_t)acc;
        else if (g == 2) P.cc_t = (uint16_t)acc;
        else P.cc_i = (uint16_t)acc;
    }
    return P;
}
constexpr TcProg kProg = tc_make_prog();

// Weight chunks in the order the producer streams them: [geometry + MLP (per view) | density head | texture fusion (per
// view) | rendering head].  The producer walks this list in a four-instruction loop; an unrolled producer (one code block
// per chunk) was 60 KB of instructions executed by one lane, which did nothing but evict the tile warps' code from the
// instruction caches.
struct TcLoadList {
    uint32_t src_off[TC_MAX_CHUNKS], bytes[TC_MAX_CHUNKS];
    int n_gm, n_q, n_t, n_i;
};
constexpr TcLoadList tc_make_loads() {
    TcLoadList L{};
    const int grp_first[4] = {ST_G1, ST_Q1, ST_T1, ST_I1}, grp_last[4] = {ST_M3, ST_Q3, ST_T4, ST_I9};
    int n = 0;
    for (int g = 0; g < 4; ++g) {
        int cnt = 0;
        for (int st = grp_first[g]; st <= grp_last[g]; ++st) {
            if (tc_step_in_regs(st)) continue;               // attention layer 2 of GeoVisFusion (and out_layer's last Linear) run in registers
            for (int c = 0; c < kProg.steps[st].nchunks; ++c, ++n, ++cnt) {
                L.src_off[n] = kProg.chunks[kProg.steps[st].chunk0 + c].src_off;
                L.bytes[n] = kProg.chunks[kProg.steps[st].chunk0 + c].bytes;
            }
        }
        if (g == 0) L.n_gm = cnt;
        else if (g == 1) L.n_q = cnt;
        else if (g == 2) L.n_t = cnt;
        else L.n_i = cnt;
    }
    return L;
}
constexpr TcLoadList kLoads = tc_make_loads();
static_assert(kLoads.n_gm == kProg.cc_gm && kLoads.n_q == kProg.cc_q && kLoads.n_t == kProg.cc_t && kLoads.n_i == kProg.cc_i,
              "producer list and issuer chunk counters disagree");
__constant__ TcLoadList c_loads = tc_make_loads();

// ================================================================================================ host: script + images
struct TcOpSpec {
    int layer, a_slot, a_col0, d_col, accum;
    std::vector<int> kmap;       // per operand column: input index of the reference layer, -1 = zero weight
    int share = 0;               // > 0: reuse the weight block of the op `share` ops earlier (same step)
};

static std::vector<int> iota_map(int from, int n, int pad_to = -1) {
    std::vector<int> m;
    for (int i = 0; i < n; ++i) m.push_back(from + i);
    while (pad_to > 0 && (int)m.size() < pad_to) m.push_back(-1);
    return m;
}

static const int kTexMap1[64] = {3, 4, 5, 6, 7, 8, 9, 10, 14, 15, 16, 17, 18, 19, 20, 21, 25, 26, 27, 28, 29, 30, 31, 32,
                                 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 51, 52, 53, 54, 55, 56, 57, 58,
                                 59, 60, 61, 62, 63, 64, 65, 66, 49, 50, 67, 68, 0, 1, 2, 11};
static const int kTexMap2[32] = {69, 70, 71, 72, 73, 74, 75, 76, 77, 78, 79, 80, 81, 82, 83, 84, 85, 86, 87, 88, 89, 90, 91, 92,
                                 12, 13, 22, 23, 24, 93, 94, 95};

static void tc_build_script(std::vector<std::vector<TcOpSpec>>& st) {
    st.assign(ST_COUNT, {});
    auto add = [&](int s, int layer, int slot, int col0, int d_col, int accum, std::vector<int> kmap, int share = 0) {
        st[s].push_back(TcOpSpec{layer, slot, col0, d_col, accum, std::move(kmap), share});
    };
    std::vector<int> tex1(kTexMap1, kTexMap1 + 64), tex2(kTexMap2, kTexMap2 + 32);
    // ---- GeoVisFusion, both scales side by side (src/networks.py:83-104)
    for (int pass = 0; pass < 2; ++pass) {                       // pass 0: attention first layer (G1), pass 1: fused first layer (G3)
        const int s = pass ? ST_G3 : ST_G1, l64 = pass ? L_GEO_F0 : L_GEO_AT0, l8 = pass ? L_GEO8_F0 : L_GEO8_AT0;
        add(s, l64, 0, 0, 0, 0, iota_map(0, 64));
        add(s, l64, 1, 0, 0, 1, iota_map(64, 64));
        add(s, l64, 2, 0, 0, 1, iota_map(128, 64));
        add(s, l64, 3, 0, 0, 1, iota_map(192, 4, 16));
        add(s, l8, 3, 16, pass ? 64 : 16, 0, iota_map(0, 28, 32));
    }
    add(ST_G2, L_GEO_AT1, 3, 48, 0, 0, iota_map(0, 10, 16));
    add(ST_G2, L_GEO8_AT1, 4, 0, 16, 0, iota_map(0, 10, 16));
    add(ST_G4, L_GEO_F1, 4, 0, 0, 0, iota_map(0, 64));
    add(ST_G4, L_GEO8_F1, 3, 48, 64, 0, iota_map(0, 8, 16));
    // ---- MLPUNet layers1 (src/utils.py:822-852); layer 0 input = [PE 294 | out64], PE in per-keypoint groups of 8
    add(ST_M0, L_MLP0, TC_M0_SLOT, 0, 0, 0, iota_map(294, 64));
    for (int s = 0; s < 6; ++s) {
        const int pslot[3] = {1, 2, TC_PE_SLOT2};
        const int step = ST_P0 + s, slot = pslot[s % 3];
        std::vector<int> m;
        for (int kl = 0; kl < (s < 5 ? 8 : 2); ++kl)
            for (int f = 0; f < 8; ++f) m.push_back(f < 7 ? f * NKPT + (8 * s + kl) : -1);
        add(step, L_MLP0, slot, 0, 0, 1, m);
    }
    add(ST_M1, L_MLP1, 1, 0, 0, 0, iota_map(0, 64));
    add(ST_M1, L_MLP1, 2, 0, 0, 1, iota_map(64, 64));
    add(ST_M2, L_MLP2, 4, 0, 0, 0, iota_map(0, 64));
    add(ST_M2, L_MLP2, 0, 0, 0, 1, iota_map(64, 64));
    add(ST_M2, L_MLP2, 3, 48, 0, 1, iota_map(128, 8, 16));
    add(ST_M3, L_MLP3, 1, 0, 0, 0, iota_map(0, 64));
    add(ST_M3, L_MLP3, 2, 0, 0, 1, iota_map(64, 56, 64));
    // ---- density head + latent compression on the pooled latent [mean | var]
    add(ST_Q1, L_POST0, 1, 0, 0, 0, iota_map(0, 64));
    add(ST_Q1, L_POST0, 2, 0, 0, 1, iota_map(64, 64));
    add(ST_Q1, L_COMPRESS, 1, 0, 64, 0, iota_map(0, 64));
    add(ST_Q1, L_COMPRESS, 2, 0, 64, 1, iota_map(64, 64));
    add(ST_Q2, L_POST1, 4, 0, 0, 0, iota_map(0, 64));
    add(ST_Q3, L_POST2, 0, 0, 0, 0, iota_map(0, 64));
    // ---- TexVisFusion (src/networks.py:281-293) + ray encoder (src/model.py:1612-1618)
    for (int pass = 0; pass < 2; ++pass) {
        const int s = pass ? ST_T3 : ST_T1, l = pass ? L_TEX_F0 : L_TEX_AT0;
        add(s, l, 1, 0, 0, 0, tex1);
        add(s, l, 2, 0, 0, 1, tex2);
    }
    add(ST_T1, L_RAY0, 3, 0, 96, 0, iota_map(0, 4, 16));
    for (int pass = 0; pass < 2; ++pass) {
        const int s = pass ? ST_T4 : ST_T2, l = pass ? L_TEX_F1 : L_TEX_AT1;
        add(s, l, 4, 0, 0, 0, iota_map(0, 64));
        add(s, l, 0, 0, 0, 1, iota_map(64, 32));
    }
    add(ST_T2, L_RAY1, 3, 16, 16, 0, iota_map(0, 16));
    // ---- IBRRenderingHead (src/model.py:1600-1636), the views batched in one step: view v works in slot 2 + v and
    // accumulator columns stride * v, views 1, 2 reuse view 0's weight blocks.  base input [mean 40 | var 40 | f 40]:
    // slot 1 = [mean 40 | var 0..23] (shared), slot 2 + v = [var 24..39 | f_v 40 | pad 8]
    for (int v = 0; v < TC_MAXV; ++v) {
        add(ST_I1, L_BASE0, 1, 0, 64 * v, 0, iota_map(0, 64), 2 * v);
        add(ST_I1, L_BASE0, 2 + v, 0, 64 * v, 1, iota_map(64, 56, 64), 2 * v);
    }
    for (int v = 0; v < TC_MAXV; ++v) add(ST_I2, L_BASE1, 2 + v, 0, 32 * v, 0, iota_map(0, 64), v);
    for (int v = 0; v < TC_MAXV; ++v) add(ST_I3, L_VIS1_0, 2 + v, 0, 48 * v, 0, iota_map(0, 32), v);
    for (int v = 0; v < TC_MAXV; ++v) add(ST_I4, L_VIS1_1, 2 + v, 32, 48 * v, 0, iota_map(0, 32), v);
    for (int v = 0; v < TC_MAXV; ++v) add(ST_I5, L_VIS2_0, 2 + v, 0, 48 * v, 0, iota_map(0, 32), v);
    for (int v = 0; v < TC_MAXV; ++v) add(ST_I6, L_VIS2_1, 2 + v, 32, 48 * v, 0, iota_map(0, 32), v);
    for (int v = 0; v < TC_MAXV; ++v) add(ST_I7, L_OUT0, 2 + v, 0, 16 * v, 0, iota_map(0, 37, 48), v);
    for (int v = 0; v < TC_MAXV; ++v) add(ST_I8, L_OUT1, 2 + v, 48, 16 * v, 0, iota_map(0, 16), v);
    for (int v = 0; v < TC_MAXV; ++v) add(ST_I9, L_OUT2, 2 + v, 0, 16 * v, 0, iota_map(0, 8, 16), v);
}

static inline uint16_t f2bf_host(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

// Builds the step / op / chunk tables and the bf16 weight images (blob) from the folded fp32 layers.
static void tc_build(const vanerf_linear* const* src_in, float ani_al, TcTables& T, TcProg& P, std::vector<uint16_t>& blob,
                     std::vector<uint16_t>* blob_lo = nullptr) {
    std::vector<std::vector<TcOpSpec>> st;
    tc_build_script(st);
#if TC_MERGE_G4
    // Fold GeoVisFusion's second fused Linear F (c x c, src/networks.py:101-104: no bias, no activation behind it) into its two
    // consumers: MLP layer 0 reads out64 in its input columns [294, 358), MLP layer 2 reads out8 in [128, 136), so
    // W[:, cols] <- W[:, cols] . F and the layers take ReLU(fused layer 1) instead (products in double, rounded once).
    const vanerf_linear* folded[L_COUNT];
    for (int l = 0; l < L_COUNT; ++l) folded[l] = src_in[l];
    std::vector<float> wf[2], bf[2];
    vanerf_linear lf[2];
    {
        const int consumer[2] = {L_MLP0, L_MLP2}, inner[2] = {L_GEO_F1, L_GEO8_F1}, col0[2] = {294, 128};
        for (int q = 0; q < 2; ++q) {
            const vanerf_linear& C = *src_in[consumer[q]];
            const vanerf_linear& F = *src_in[inner[q]];
            const int c = F.out_dim;
            wf[q].assign(C.w, C.w + (size_t)C.out_dim * C.in_dim);
            if (C.b) bf[q].assign(C.b, C.b + C.out_dim);
            for (int n = 0; n < C.out_dim; ++n) {
                for (int j = 0; j < F.in_dim; ++j) {
                    double acc = 0.0;
                    for (int i = 0; i < c; ++i) acc += (double)C.w[(size_t)n * C.in_dim + col0[q] + i] * (double)F.w[(size_t)i * F.in_dim + j];
                    wf[q][(size_t)n * C.in_dim + col0[q] + j] = (float)acc;
                }
                if (F.b) {             // not the case in the reference (bias=False); kept exact if a checkpoint has one
                    double acc = C.b ? (double)C.b[n] : 0.0;
                    for (int i = 0; i < c; ++i) acc += (double)C.w[(size_t)n * C.in_dim + col0[q] + i] * (double)F.b[i];
                    if (bf[q].empty()) bf[q].assign(C.out_dim, 0.0f);
                    bf[q][n] = (float)acc;
                }
            }
            lf[q] = C;
            lf[q].w = wf[q].data();
            lf[q].b = bf[q].empty() ? nullptr : bf[q].data();
            folded[consumer[q]] = &lf[q];
        }
    }
    const vanerf_linear* const* src = folded;
#else
    const vanerf_linear* const* src = src_in;
#endif
    memset(&T, 0, sizeof(T));
    memset(&P, 0, sizeof(P));
    blob.clear();
    if (blob_lo) blob_lo->clear();
    int n_ops = 0, n_chunks = 0;
    for (int s = 0; s < ST_COUNT; ++s) {
        P.steps[s].op0 = (uint16_t)n_ops;
        P.steps[s].chunk0 = (uint16_t)n_chunks;
        uint32_t cur_bytes = 0;
        int rel = -1;
        for (const TcOpSpec& o : st[s]) {
            const vanerf_linear& L = *src[o.layer];
            const int n_pad = (L.out_dim + 15) & ~15;
            const int ncols = (int)o.kmap.size();
            const uint32_t bytes = (uint32_t)n_pad * 128;
            if (!o.share && (rel < 0 || cur_bytes + bytes > TC_SLOT)) {          // open a new chunk
                if (rel >= 0) P.ops[n_ops - 1].last_in_chunk = 1;
                ++rel;
                P.chunks[n_chunks].src_off = (uint32_t)(blob.size() * 2);
                P.chunks[n_chunks].bytes = 0;
                ++n_chunks;
                cur_bytes = 0;
            }
            TcOp& d = P.ops[n_ops++];
            d.a_off = (uint32_t)(o.a_slot * TC_SLOT + (o.a_col0 / 16) * 32);
            d.b_off = o.share ? P.ops[n_ops - 1 - o.share].b_off : cur_bytes;
            d.idesc = tc::umma_idesc_bf16(128, n_pad);
            d.d_col = (uint16_t)o.d_col;
            d.nk = (uint8_t)(ncols / 16);
            d.accum = (uint8_t)o.accum;
            d.chunk_rel = o.share ? P.ops[n_ops - 1 - o.share].chunk_rel : (uint8_t)rel;
            d.last_in_chunk = 0;
            if (o.share) continue;                                               // weights already in the chunk
            const size_t base = blob.size();
            blob.resize(base + bytes / 2, 0);
            if (blob_lo) blob_lo->resize(base + bytes / 2, 0);
            for (int n = 0; n < L.out_dim; ++n)
                for (int k = 0; k < ncols; ++k) {
                    const int ki = o.kmap[k];
                    if (ki < 0) continue;
                    const size_t byte = (size_t)(n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 3) ^ n) & 7) << 4) + (k & 7) * 2;
                    const float wv = L.w[(size_t)n * L.in_dim + ki];
                    const uint16_t hi = f2bf_host(wv);
                    blob[base + byte / 2] = hi;
                    if (blob_lo) {                 // split path: w = hi + lo to 16 mantissa bits
                        const uint32_t hb = (uint32_t)hi << 16;
                        float hf;
                        memcpy(&hf, &hb, 4);
                        (*blob_lo)[base + byte / 2] = f2bf_host(wv - hf);
                    }
                }
            cur_bytes += bytes;
            P.chunks[n_chunks - 1].bytes = cur_bytes;
        }
        if (n_ops > P.steps[s].op0) P.ops[n_ops - 1].last_in_chunk = 1;
        P.steps[s].nops = (uint16_t)(n_ops - P.steps[s].op0);
        P.steps[s].nchunks = (uint16_t)(n_chunks - P.steps[s].chunk0);
    }
    int boff = 0;
    for (int l = 0; l < L_COUNT; ++l) {
        T.bias_off[l] = (uint16_t)boff;
        if (src[l]->b) {
            for (int n = 0; n < src[l]->out_dim; ++n) T.bias[boff + n] = src[l]->b[n];
            boff += (src[l]->out_dim + 15) & ~15;
        }
    }
    T.bias_off[L_COUNT] = (uint16_t)boff;
    {   // chunk offsets of the steps inside their group (attention layer 2 of GeoVisFusion runs in registers: no ST_G2)
        const int grp_first[4] = {ST_G1, ST_Q1, ST_T1, ST_I1}, grp_last[4] = {ST_M3, ST_Q3, ST_T4, ST_I9};
        uint16_t* tot[4] = {&P.cc_gm, &P.cc_q, &P.cc_t, &P.cc_i};
        for (int g = 0; g < 4; ++g) {
            int acc = 0;
            for (int st = grp_first[g]; st <= grp_last[g]; ++st) {
                P.cc_off[st] = (uint16_t)acc;
                if (!tc_step_in_regs(st)) acc += P.steps[st].nchunks;
            }
            *tot[g] = (uint16_t)acc;
        }
    }
    T.ani_al_abs = fabsf(ani_al);
    for (int i = 0; i < 8; ++i) T.out2w[i] = src[L_OUT2]->w[i];
    for (int sc = 0; sc < 2; ++sc) {
        const vanerf_linear& L = *src[sc ? L_GEO8_AT1 : L_GEO_AT1];
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 10; ++i) T.at2[sc * 36 + j * 12 + i] = L.w[(size_t)j * L.in_dim + i];
    }
}

// true when the tables the packing script produced are the ones the kernels were compiled with
static bool tc_program_matches(const TcProg& P) {
    static const TcProg K = kProg;
    for (int s = 0; s < ST_COUNT; ++s)
        if (P.steps[s].op0 != K.steps[s].op0 || P.steps[s].nops != K.steps[s].nops || P.steps[s].chunk0 != K.steps[s].chunk0 ||
            P.steps[s].nchunks != K.steps[s].nchunks || P.cc_off[s] != K.cc_off[s]) return false;
    for (int i = 0; i < TC_MAX_OPS; ++i)
        if (P.ops[i].a_off != K.ops[i].a_off || P.ops[i].b_off != K.ops[i].b_off || P.ops[i].idesc != K.ops[i].idesc ||
            P.ops[i].d_col != K.ops[i].d_col || P.ops[i].nk != K.ops[i].nk || P.ops[i].accum != K.ops[i].accum ||
            P.ops[i].chunk_rel != K.ops[i].chunk_rel || P.ops[i].last_in_chunk != K.ops[i].last_in_chunk) return false;
    for (int i = 0; i < TC_MAX_CHUNKS; ++i)
        if (P.chunks[i].src_off != K.chunks[i].src_off || P.chunks[i].bytes != K.chunks[i].bytes) return false;
    return P.cc_gm == K.cc_gm && P.cc_q == K.cc_q && P.cc_t == K.cc_t && P.cc_i == K.cc_i;
}

static_assert(TC_TAB_PARAM || sizeof(TcTables) <= TC_TAB_BYTES, "TC_TAB_BYTES too small");
static_assert(sizeof(TcTables) <= 16384, "TcTables travels as a kernel parameter (limit 32 764 bytes together with TcArgs)");
static_assert(sizeof(TcProg) <= 8192, "TcProg must stay a small part of constant memory");
// ================================================================================================ device
// Optional cycle trace of CTA 0 / thread 0 (vanerf_tc_profile), compiled in only with -DVANERF_TC_TRACE: entries
// (tag << 48 | clock64 & (2^48-1)) collected in shared memory and flushed to d_tc_prof when the kernel ends.
extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
#ifdef VANERF_TC_TRACE
__device__ long long* d_tc_prof = nullptr;
__device__ int d_tc_prof_cap = 0;
__device__ int d_tc_prof_n = 0;
// traced threads of CTA 0: thread 0 (tile 0, row 0) and lane 0 of tile 0's MMA issuer warp (tags + 100000)
#define TC_TRACE_ISSUER (TC_TILES * TC_EPI_THREADS + 32)
__device__ __forceinline__ void tc_prof(int tag) {
    if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == TC_TRACE_ISSUER)) {
        const int who = threadIdx.x == 0 ? 0 : 1;
        int* ctr = reinterpret_cast<int*>(tc_smem_raw + TC_OFF_CTRL + 496 + 4 * who);
        const int i = *ctr;
        if (i < 0 || i >= TC_TRACE_N / 2) return;       // -1 = tracing off
        reinterpret_cast<unsigned long long*>(tc_smem_raw + TC_OFF_TRACE)[who * (TC_TRACE_N / 2) + i] =
            ((unsigned long long)(tag + 100000 * who) << 40) | ((unsigned long long)clock64() & 0xffffffffffull);
        *ctr = i + 1;
    }
}
#define TC_PROF(tag) tc_prof(tag)
#else
#define TC_PROF(tag) ((void)0)
#endif

struct TcShared {                      // control block behind the tables
    uint64_t wfull[TC_NRING], wempty[TC_NRING], acc_bar[TC_TILES], rec_bar[TC_TILES], pfree[TC_TILES][3];
    uint64_t ready[TC_TILES][TC_NREADY];   // operands of step n published (8 arrivals: one per tile warp), n mod TC_NREADY
    uint32_t tmem_base;
    int abort_flag[8];                 // [0] first code that gave up, [1 + code/100] pending wait classes (tc_prims.cuh)
    int stop[TC_TILES];                // per tile group: leave the persistent loop (abort seen by the group's leader)
};

struct TcArgs {
    const TcTables* tab;               // global copy of the tables
    const unsigned char* wblob;
    const unsigned char* rec;          // (n_tiles, V, 5, 16 KB)
    const unsigned char* aux;          // (n_tiles*128, V, 64)
    int V, n_chunk;
    long long sample0;
    unsigned wblob_lo_off;             // split path: byte offset of the lo weight images inside wblob
    float* rgba;                       // (N,5) or NULL
    float* raw_out;                    // (N,5) or NULL
    float* dbg_latent;                 // (N,128) or NULL
    int* err;                          // mapped host int: nonzero = a bounded wait gave up (code)
};

enum TcAct { TA_NONE = 0, TA_RELU, TA_SOFTPLUS, TA_SIGMOID, TA_ELU };
__device__ __forceinline__ float tc_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <int ACT> __device__ __forceinline__ float tc_act(float x) {
    if (ACT == TA_RELU) return fmaxf(x, 0.0f);
    if (ACT == TA_SOFTPLUS) return fmaxf(x, 0.0f) + 0.01f * __logf(1.0f + __expf(-100.0f * fabsf(x)));   // Softplus(beta=100)
#if TC_ABLATE & 8
    if (ACT == TA_SIGMOID || ACT == TA_ELU) return x;
#endif
    if (ACT == TA_SIGMOID) return __fdividef(1.0f, 1.0f + tc_ex2(-1.44269504f * x));
    if (ACT == TA_ELU) return x > 0.0f ? x : tc_ex2(1.44269504f * x) - 1.0f;
    return x;
}
// Softplus(beta = 100, threshold 20) of two values, evaluated on packed bf16 (the result is rounded to bf16 anyway):
//   softplus(x) = max(x, 0) + log1p(exp(-100 |x|)) / 100,   w = exp(-100 |x|) in (0, 1] by one packed MUFU ex2,
//   log1p(w) / 100 ~ w (c0 + w (c1 + w c2))  (least squares on [0, 1], |error| < 7e-6 in the output).
// Above the reference's threshold (100 x > 20) the correction is < 2.1e-11, i.e. the reference's linear branch.
__device__ __forceinline__ uint32_t tc_softplus2(float a, float b) {
    uint32_t h, na, w, p, r;
#if TC_ABLATE & 1
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
#endif
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
    na = h | 0x80008000u;                                                        // -|h|
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(w) : "r"(na), "r"(0x43104310u));      // * 144 (100 log2 e)
    asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(w) : "r"(w));
    asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(w), "r"(0x3A933A93u), "r"(0xBB85BB85u));   // c2 = 1.1234e-3, c1 = -4.0516e-3
    asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(p), "r"(w), "r"(0x3C223C22u));              // c0 = 9.8642e-3
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(p) : "r"(p), "r"(w));
    asm("fma.rn.relu.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(h), "r"(0x3F803F80u), "r"(0u));        // max(h, 0)
    asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(r), "r"(p));
    return r;
}
template <int ACT> __device__ __forceinline__ uint32_t tc_pack_act(float a, float b) {    // act(a) -> low half, act(b) -> high half
    if (ACT == TA_SOFTPLUS) return tc_softplus2(a, b);
    if (ACT == TA_RELU) {
        uint32_t d;
        asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
        return d;
    }
    return tc::pack_bf16(tc_act<ACT>(a), tc_act<ACT>(b));
}
__device__ __forceinline__ uint32_t tc_mul2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t tc_dup_bf16(float g) { return tc::pack_bf16(g, g); }
// (a * s, b * s) -> packed bf16; with TC_F32X2 the two products are one FMUL2
__device__ __forceinline__ uint32_t tc_pack_scaled(float a, float b, float s) {
#if TC_F32X2
    const float2 p = __fmul2_rn(make_float2(a, b), make_float2(s, s));
    return tc::pack_bf16(p.x, p.y);
#else
    return tc::pack_bf16(a * s, b * s);
#endif
}

__device__ __forceinline__ float tc_sel3(const float (&a)[3], int v) { return v == 0 ? a[0] : (v == 1 ? a[1] : a[2]); }
__device__ __forceinline__ void tc_set3(float (&a)[3], int v, float x) {
    a[0] = v == 0 ? x : a[0]; a[1] = v == 1 ? x : a[1]; a[2] = v == 2 ? x : a[2];
}
__device__ __forceinline__ bool tc_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag, int code) {
    return tc::mbar_wait(bar, parity, abort_flag, code);
}

__constant__ TcProg c_prog;

__device__ __forceinline__ bool tc_elect() {            // one lane of the (converged) warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// Shared-memory matrix descriptor of the K-major 128B-swizzle layout, split: the high word is constant.
#define TC_DESC_HI ((uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29))
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr) {
    return ((uint64_t)TC_DESC_HI << 32) | (uint64_t)(((smem_addr & 0x3FFFFu) >> 4) | (1u << 16));
}

// Issues op I (and the following ones) of step ST: everything but the ring slot of the weight chunk is an immediate.
// Whole warp, converged; `lead` = the elected lane that executes the MMAs and commits.
template <int ST, int I, bool SPLIT>
__device__ __forceinline__ void tc_issue_op(uint32_t cc, uint64_t bd_slot, uint32_t slot_i, uint64_t bd_slot_lo, uint32_t slot_lo, TcShared* sh,
                                            uint64_t ad_base, uint64_t bd_base, uint32_t tmem, bool lead) {
    constexpr TcStep S = kProg.steps[ST];
    constexpr TcOp op = kProg.ops[S.op0 + I];
    constexpr bool first_in_chunk = I == 0 || kProg.ops[S.op0 + (I > 0 ? I - 1 : 0)].last_in_chunk != 0;
    if (first_in_chunk) {
        const uint32_t c = (SPLIT ? 2u : 1u) * (cc + op.chunk_rel);       // split: logical chunk c = ring chunks 2c (hi), 2c + 1 (lo)
        slot_i = c % TC_NRING;
        bd_slot = bd_base + (uint64_t)(slot_i * (TC_SLOT >> 4));
        if (SPLIT) {
            slot_lo = (c + 1) % TC_NRING;
            bd_slot_lo = bd_base + (uint64_t)(slot_lo * (TC_SLOT >> 4));
        }
#if TC_PREWAIT
        if (SPLIT)
#endif
        {
            TC_PROF(7000);
            tc::mbar_wait(&sh->wfull[slot_i], (c / TC_NRING) & 1, sh->abort_flag, 100 + ST);
            if (SPLIT) tc::mbar_wait(&sh->wfull[slot_lo], ((c + 1) / TC_NRING) & 1, sh->abort_flag, 100 + ST);
            tc::tcgen05_fence_after();
            TC_PROF(7100);
        }
    }
    // descriptors = base descriptor + (byte offset >> 4): the start-address field (14 bits of address >> 4) cannot
    // carry into its neighbours because every operand lies inside the CTA's 227 KB of shared memory
    const uint64_t ad = ad_base + (uint64_t)(op.a_off >> 4);
    const uint64_t bd = bd_slot + (uint64_t)(op.b_off >> 4);
    if (lead) {
#pragma unroll
        for (int k = 0; k < ((TC_ABLATE & 4) ? 1 : op.nk); ++k) {            // +32 bytes (16 bf16) per K step inside the 128-byte swizzled row
            tc::umma_bf16(tmem + op.d_col, ad + 2 * k, bd + 2 * k, op.idesc, (op.accum || k > 0) ? 1u : 0u);
            if (SPLIT) {
                tc::umma_bf16(tmem + op.d_col, ad + (uint64_t)(TC_LO_OFF >> 4) + 2 * k, bd + 2 * k, op.idesc, 1u);                    // lo(A) x hi(W)
                tc::umma_bf16(tmem + op.d_col, ad + 2 * k, bd_slot_lo + (uint64_t)(op.b_off >> 4) + 2 * k, op.idesc, 1u);             // hi(A) x lo(W)
            }
        }
        if (op.last_in_chunk) {
            tc::umma_commit(&sh->wempty[slot_i]);
            if (SPLIT) tc::umma_commit(&sh->wempty[slot_lo]);
        }
    }
    if constexpr (I + 1 < S.nops) tc_issue_op<ST, I + 1, SPLIT>(cc, bd_slot, slot_i, bd_slot_lo, slot_lo, sh, ad_base, bd_base, tmem, lead);
}
// Waits for every weight chunk of step ST (TC_PREWAIT): done BEFORE the wait for the step's operands, because the weights
// are streamed a step ahead and have normally landed long before the tile's epilogue publishes; the ~100-cycle
// already-complete try_wait per chunk then leaves the tile's critical path.  No deadlock: a chunk's ring slot only
// depends on MMAs of earlier steps (issued by this warp in program order, by the other tile's issuer at its own pace).
template <int ST, int I>
__device__ __forceinline__ void tc_prewait_chunks(uint32_t cc, TcShared* sh) {
    constexpr TcStep S = kProg.steps[ST];
    constexpr TcOp op = kProg.ops[S.op0 + I];
    constexpr bool first_in_chunk = I == 0 || kProg.ops[S.op0 + (I > 0 ? I - 1 : 0)].last_in_chunk != 0;
    if (first_in_chunk) {
        const uint32_t c = cc + op.chunk_rel;
#if !(TC_ABLATE & 64)
        tc::mbar_wait(&sh->wfull[c % TC_NRING], (c / TC_NRING) & 1, sh->abort_flag, 100 + ST);
#endif
    }
    if constexpr (I + 1 < S.nops) tc_prewait_chunks<ST, I + 1>(cc, sh);
}
// One step of an MMA issuer warp: wait until the tile's eight warps have published the operands of step number n
// (and finished reading the accumulators the step overwrites), then issue the step's MMAs and commit.
// COMMIT: 0 = accumulator barrier, 1..3 = PE ring-slot barrier (COMMIT - 1), -1 = none.
// cc_base = index of the first weight chunk of this iteration of the step's group (see TcProg::cc_off).
template <int ST, int COMMIT, bool SPLIT>
__device__ __forceinline__ void tc_issuer_step(uint32_t& n, uint32_t cc_base, TcShared* sh, uint32_t act_u32, uint32_t ring_u32,
                                               uint32_t tmem, int tg, bool lead) {
#if TC_PREWAIT
    if constexpr (!SPLIT) {              // split path: a step's (hi, lo) chunk pairs can exceed the ring, no pre-wait
        constexpr uint32_t cc_off_pre = kProg.cc_off[ST];
        tc_prewait_chunks<ST, 0>(cc_base + cc_off_pre, sh);
    }
#endif
    TC_PROF(8000 + ST);                  // issuer: previous step issued, waiting for this step's operands
    const bool ok = tc::mbar_wait(&sh->ready[tg][n % TC_NREADY], (n / TC_NREADY) & 1, sh->abort_flag, 600 + ST);
    ++n;
    tc::tcgen05_fence_after();
    TC_PROF(9000 + ST);                  // issuer: operands ready
    constexpr uint32_t cc_off = kProg.cc_off[ST];
    // The operand addresses are tied to the outcome of the wait (+0 unless the wait was abandoned, in which case the
    // kernel is draining and results are discarded): otherwise the compiler computes the descriptors of all 140-odd
    // MMAs ahead of the waits and spills them to local memory.
    const uint32_t never = ok ? 0u : 16u;
    tc_issue_op<ST, 0, SPLIT>(cc_base + cc_off, 0, 0, 0, 0, sh, tc_desc(act_u32 + never), tc_desc(ring_u32 + never), tmem, lead);
    if (lead) {
        if (COMMIT == 0) tc::umma_commit(&sh->acc_bar[tg]);
        else if (COMMIT > 0) tc::umma_commit(&sh->pfree[tg][COMMIT > 0 ? COMMIT - 1 : 0]);
    }
    __syncwarp();
    TC_PROF(3000 + ST);                  // MMAs issued (includes the wait for the weight chunk)
}

// Table-driven variant for the MMA self test (one step of arbitrary K, N read from __constant__ c_prog).
__device__ __forceinline__ void tc_issue_ops_dyn(int st, uint32_t cc_base, TcShared* sh, uint32_t act_u32, uint32_t ring_u32, uint32_t tmem, int tg) {
    const TcStep S = c_prog.steps[st];
    const uint32_t cc = cc_base + c_prog.cc_off[st];
    const bool lead = tc_elect();
    uint32_t slot_i = 0;
    int cur = -1;
#pragma unroll 1
    for (int i = 0; i < S.nops; ++i) {
        const TcOp op = c_prog.ops[S.op0 + i];
        if ((int)op.chunk_rel != cur) {
            cur = op.chunk_rel;
            const uint32_t c = cc + op.chunk_rel;
            slot_i = c % TC_NRING;
            tc::mbar_wait(&sh->wfull[slot_i], (c / TC_NRING) & 1, sh->abort_flag, 100 + st);
            tc::tcgen05_fence_after();
        }
        const uint64_t ad = tc_desc(act_u32 + op.a_off);
        const uint64_t bd = tc_desc(ring_u32 + slot_i * TC_SLOT + op.b_off);
        if (lead) {
#pragma unroll 1
            for (int k = 0; k < op.nk; ++k)
                tc::umma_bf16(tmem + op.d_col, ad + 2 * k, bd + 2 * k, op.idesc, (op.accum || k > 0) ? 1u : 0u);
            if (op.last_in_chunk) tc::umma_commit(&sh->wempty[slot_i]);
        }
    }
    if (lead) tc::umma_commit(&sh->acc_bar[tg]);
    __syncwarp();
}

struct TcTile {
    unsigned char* smem;      // dynamic smem base
    unsigned char* act;       // operand slots of this tile
    TcShared* sh;
    const TcTables* tb;       // tables in shared memory
    uint32_t trow;            // TMEM address of this thread's row, first column of this tile's half
    uint32_t row_off, rx;     // byte offset of this thread's row inside a slot, row & 7 (swizzle key)
    int row, half, tg;
    uint32_t acc_phase, rec_phase, pfree_bits;           // pfree_bits: bit ps = phase parity of PE ring slot ps
    uint32_t step_n;                                     // number of steps signalled so far (selects the operand-ready barrier)

    __device__ __forceinline__ unsigned char* slot(int s) const { return act + s * TC_SLOT; }
    __device__ __forceinline__ uint32_t coff(int chunk) const { return row_off + (((uint32_t)chunk ^ rx) << 4); }
    __device__ __forceinline__ void st_chunk(int s, int chunk, const uint4& q) const {
        *reinterpret_cast<uint4*>(slot(s) + coff(chunk)) = q;
    }
    __device__ __forceinline__ void st_chunk_lo(int s, int chunk, const uint4& q) const {
        *reinterpret_cast<uint4*>(slot(s) + TC_LO_OFF + coff(chunk)) = q;
    }
    __device__ __forceinline__ uint4 ld_chunk(int s, int chunk) const {
        return *reinterpret_cast<const uint4*>(slot(s) + coff(chunk));
    }
    // ---- split-aware operand access: 8 fp32 values <-> one 16-byte chunk (hi image) [+ the same chunk of the lo image]
    template <bool SPLIT>
    __device__ __forceinline__ void put8(int s, int chunk, const float (&f)[8]) const {
        const uint4 hi = pack8(f);
        st_chunk(s, chunk, hi);
        if (SPLIT) {
            float fh[8], lo[8];
            unpack8(hi, fh);
#pragma unroll
            for (int i = 0; i < 8; ++i) lo[i] = f[i] - fh[i];
            *reinterpret_cast<uint4*>(slot(s) + TC_LO_OFF + coff(chunk)) = pack8(lo);
        }
    }
    template <bool SPLIT>
    __device__ __forceinline__ void put8(int s, int chunk, float a, float b, float c, float d, float e, float f, float g, float h) const {
        const float v[8] = {a, b, c, d, e, f, g, h};
        put8<SPLIT>(s, chunk, v);
    }
    template <bool SPLIT>
    __device__ __forceinline__ void get8(int s, int chunk, float (&f)[8]) const {
        unpack8(ld_chunk(s, chunk), f);
        if (SPLIT) {
            float lo[8];
            unpack8(*reinterpret_cast<const uint4*>(slot(s) + TC_LO_OFF + coff(chunk)), lo);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] += lo[i];
        }
    }
    template <bool SPLIT>
    __device__ __forceinline__ void zero8(int s, int chunk) const {
        st_chunk(s, chunk, make_uint4(0, 0, 0, 0));
        if (SPLIT) *reinterpret_cast<uint4*>(slot(s) + TC_LO_OFF + coff(chunk)) = make_uint4(0, 0, 0, 0);
    }
    __device__ __forceinline__ void ld16(int col, float (&v)[16]) const {
        uint32_t r[16];
        tc::tmem_ld16(trow + col, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
    }
    __device__ __forceinline__ void ld8(int col, float (&v)[8]) const {
        uint32_t r[8];
        tc::tmem_ld8(trow + col, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
    }
    __device__ __forceinline__ void st16(int col, const float (&v)[16]) const {
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(v[i]);
        tc::tmem_st16(trow + col, r);
    }
    __device__ __forceinline__ void st8(int col, const float (&v)[8]) const {
        uint32_t r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(v[i]);
        tc::tmem_st8(trow + col, r);
    }
    // Publishes the calling warp's operand writes (and orders its TMEM reads before the coming MMAs) and tells the
    // tile's MMA issuer warp: one arrival per warp on the operand-ready barrier of this step number.  Non-blocking; a
    // warp can be at most TC_NREADY - 1 steps ahead of the slowest warp of its tile (the PE steps have no accumulator
    // wait in between, their ring-slot waits bound the lead to 3), so TC_NREADY barriers are never lapped.
    __device__ __forceinline__ void issue(int st, int commit_to) {
        (void)st; (void)commit_to;            // the issuer warp walks the same static sequence (TcProg::seq_*)
        TC_PROF(1000 + st);                   // epilogue of the previous step done (this thread)
#if !(TC_ABLATE & 32)
        tc::fence_proxy_async();
#endif
        TC_PROF(2000 + st);                   // proxy fence done
        tc::tcgen05_fence_before();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) tc::mbar_arrive(&sh->ready[tg][step_n % TC_NREADY]);
        ++step_n;
    }
    __device__ __forceinline__ void wait_acc(int st) {
        tc_wait(&sh->acc_bar[tg], acc_phase, sh->abort_flag, 200 + st);
        acc_phase ^= 1;
        tc::tcgen05_fence_after();
        TC_PROF(4000 + st);              // accumulator complete
    }
    __device__ __forceinline__ void step(int st) { issue(st, 0); wait_acc(st); }
    // multiply the 8 values of a chunk by per-element gates (fp32 product, one rounding back to bf16 [hi + lo])
    template <bool SPLIT>
    __device__ __forceinline__ void gate_chunk(int s, int chunk, const float (&g)[8]) const {
        float f[8];
        get8<SPLIT>(s, chunk, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] *= g[i];
        put8<SPLIT>(s, chunk, f);
    }
    template <bool SPLIT>
    __device__ __forceinline__ void gate_chunk1(int s, int chunk, float g) const {
#if TC_ABLATE & 16
        return;
#endif
        float f[8];
        get8<SPLIT>(s, chunk, f);
#if TC_F32X2
        const float2 g2 = make_float2(g, g);
#pragma unroll
        for (int i = 0; i < 8; i += 2) { const float2 p = __fmul2_rn(make_float2(f[i], f[i + 1]), g2); f[i] = p.x; f[i + 1] = p.y; }
#else
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] *= g;
#endif
        put8<SPLIT>(s, chunk, f);
    }
};

// acc[16 columns in r] + bias -> ACT -> bf16 -> operand chunks `chunk`, `chunk + 1` of the slot at `slot_base`.
// SPLIT: the activation runs in fp32 and the result is stored as bf16 hi + bf16 lo.
template <int ACT, bool SPLIT>
__device__ __forceinline__ void tc_epi_group(const uint32_t (&r)[16], const float* bias, unsigned char* slot_base, const TcTile& t, int chunk) {
    float v[16];
    if (bias) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 b4 = reinterpret_cast<const float4*>(bias)[i];      // constant bank, warp-uniform address
#if TC_F32X2
            const float2 lo = __fadd2_rn(make_float2(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])), make_float2(b4.x, b4.y));
            const float2 hi = __fadd2_rn(make_float2(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])), make_float2(b4.z, b4.w));
            v[4 * i] = lo.x; v[4 * i + 1] = lo.y; v[4 * i + 2] = hi.x; v[4 * i + 3] = hi.y;
#else
            v[4 * i] = __uint_as_float(r[4 * i]) + b4.x; v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + b4.y;
            v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + b4.z; v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + b4.w;
#endif
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
    }
    if constexpr (SPLIT) {
        float a[8], b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] = tc_act<ACT>(v[i]); b[i] = tc_act<ACT>(v[8 + i]); }
        const uint4 ha = pack8(a), hb = pack8(b);
        float fa[8], fb[8];
        unpack8(ha, fa); unpack8(hb, fb);
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] -= fa[i]; b[i] -= fb[i]; }
        *reinterpret_cast<uint4*>(slot_base + t.coff(chunk)) = ha;
        *reinterpret_cast<uint4*>(slot_base + t.coff(chunk + 1)) = hb;
        *reinterpret_cast<uint4*>(slot_base + TC_LO_OFF + t.coff(chunk)) = pack8(a);
        *reinterpret_cast<uint4*>(slot_base + TC_LO_OFF + t.coff(chunk + 1)) = pack8(b);
    } else {
        *reinterpret_cast<uint4*>(slot_base + t.coff(chunk)) =
            make_uint4(tc_pack_act<ACT>(v[0], v[1]), tc_pack_act<ACT>(v[2], v[3]), tc_pack_act<ACT>(v[4], v[5]), tc_pack_act<ACT>(v[6], v[7]));
        *reinterpret_cast<uint4*>(slot_base + t.coff(chunk + 1)) =
            make_uint4(tc_pack_act<ACT>(v[8], v[9]), tc_pack_act<ACT>(v[10], v[11]), tc_pack_act<ACT>(v[12], v[13]), tc_pack_act<ACT>(v[14], v[15]));
    }
}
// acc[col0 + 16 g .. +16) -> chunks (chunk0 + 2 g, +1), g < N16.  The TMEM load of group g + 1 is in flight while
// group g is processed.
template <int ACT, int N16, bool SPLIT = false>
__device__ __forceinline__ void tc_epi_store(const TcTile& t, int col0, const float* bias, unsigned char* slot_base, int chunk0) {
    uint32_t ra[16], rb[16];                 // double buffer: a buffer is only read after the wait that follows its load
    tc::tmem_ld16(t.trow + col0, ra);
#pragma unroll
    for (int g = 0; g < N16; g += 2) {
        tc::tmem_ld_wait();
        if (g + 1 < N16) tc::tmem_ld16(t.trow + col0 + 16 * (g + 1), rb);
        tc_epi_group<ACT, SPLIT>(ra, bias ? bias + 16 * g : nullptr, slot_base, t, chunk0 + 2 * g);
        if (g + 1 < N16) {
            tc::tmem_ld_wait();
            if (g + 2 < N16) tc::tmem_ld16(t.trow + col0 + 16 * (g + 2), ra);
            tc_epi_group<ACT, SPLIT>(rb, bias ? bias + 16 * (g + 1) : nullptr, slot_base, t, chunk0 + 2 * (g + 1));
        }
    }
}
#define EPI(ACT, col0, n16, bias, s, chunk0) tc_epi_store<ACT, n16, SPLIT>(t, (col0), (bias), t.slot(s), (chunk0))
#define BIASP(l) (t.tb->bias + t.tb->bias_off[l])          // bias offsets are multiples of 16 floats
#define NOBIAS ((const float*)nullptr)


// n_consumers = tile groups that issue MMAs (and therefore release ring slots)
__device__ __forceinline__ void tc_setup(unsigned char* smem, TcShared* sh, const TcTables* tab_g, int n_consumers) {
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < TC_NRING; ++i) { tc::mbar_init(&sh->wfull[i], 1); tc::mbar_init(&sh->wempty[i], n_consumers); }
        for (int g = 0; g < TC_TILES; ++g) {
            tc::mbar_init(&sh->acc_bar[g], 1);
            tc::mbar_init(&sh->rec_bar[g], 1);
            for (int i = 0; i < 3; ++i) tc::mbar_init(&sh->pfree[g][i], 1);
            for (int i = 0; i < TC_NREADY; ++i) tc::mbar_init(&sh->ready[g][i], TC_EPI_THREADS / 32);
        }
        for (int i = 0; i < 8; ++i) sh->abort_flag[i] = 0;
        for (int g = 0; g < TC_TILES; ++g) sh->stop[g] = 0;
        if ((tc::smem_u32(smem) & 1023u) != 0) sh->abort_flag[0] = 1;       // operand slots need 1024-byte alignment
#ifdef VANERF_TC_TRACE
        for (int w = 0; w < 2; ++w)
            *reinterpret_cast<int*>(smem + TC_OFF_CTRL + 496 + 4 * w) = (blockIdx.x == 0 && d_tc_prof != nullptr) ? 0 : -1;
#endif
        tc::mbar_fence_init();
    }
#if !TC_TAB_PARAM
    if (tab_g) {   // tables -> shared memory
        const uint4* src = reinterpret_cast<const uint4*>(tab_g);
        uint4* dst = reinterpret_cast<uint4*>(smem + TC_OFF_TAB);
        for (int i = tid; i < (int)(sizeof(TcTables) / 16); i += blockDim.x) dst[i] = src[i];
    }
#else
    (void)tab_g;
#endif
    if (warp == 0) tc::tmem_alloc(&sh->tmem_base, TC_TMEM_COLS);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
}
__device__ __forceinline__ void tc_teardown(TcShared* sh, int* err) {
    const int tid = threadIdx.x, warp = tid >> 5;
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    if (tid == 0 && sh->abort_flag[0] && atomicCAS(err, 0, sh->abort_flag[0]) == 0)
        for (int i = 1; i < 8; ++i) err[i] = sh->abort_flag[i];
#ifdef VANERF_TC_TRACE
    if (tid == 0 && blockIdx.x == 0 && d_tc_prof) {             // flush the cycle traces
        int base = d_tc_prof_n;
        for (int w = 0; w < 2; ++w) {
            const int n = *reinterpret_cast<int*>(tc_smem_raw + TC_OFF_CTRL + 496 + 4 * w);
            for (int i = 0; i < n && base < d_tc_prof_cap; ++i, ++base) {
                const unsigned long long e = reinterpret_cast<unsigned long long*>(tc_smem_raw + TC_OFF_TRACE)[w * (TC_TRACE_N / 2) + i];
                d_tc_prof[2 * base] = (long long)(e >> 40);
                d_tc_prof[2 * base + 1] = (long long)(e & 0xffffffffffull);
            }
        }
        d_tc_prof_n = base;
    }
#endif
    if (warp == 0) tc::tmem_dealloc(sh->tmem_base, TC_TMEM_COLS);
}
__device__ __forceinline__ void tc_tile_init(TcTile& t, unsigned char* smem, TcShared* sh, int lane, const TcTables* tab_param = nullptr) {
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // warp-uniform by construction
    t.tg = warp >> 3;
    t.smem = smem;
    t.act = smem + t.tg * (TC_NACT * TC_SLOT);
    t.sh = sh;
#if TC_TAB_PARAM
    t.tb = tab_param;
#else
    (void)tab_param;
    t.tb = reinterpret_cast<const TcTables*>(smem + TC_OFF_TAB);
#endif
    t.row = 32 * (warp & 3) + lane;
    t.row_off = (uint32_t)((t.row >> 3) * 1024 + (t.row & 7) * 128);
    t.rx = (uint32_t)(t.row & 7);
    t.half = (warp >> 2) & 1;
    t.trow = sh->tmem_base + t.tg * TC_TMEM_TILE + ((uint32_t)(32 * (warp & 3)) << 16);
    t.acc_phase = 0; t.rec_phase = 0; t.step_n = 0;
    t.pfree_bits = 0;
}
__device__ __forceinline__ void tc_load_chunk(uint32_t src_off, uint32_t bytes, uint32_t& cc, unsigned char* smem, const unsigned char* wblob, int code) {
    TcShared* sh = reinterpret_cast<TcShared*>(smem + TC_OFF_CTRL);
    const uint32_t s = cc % TC_NRING;
    tc::mbar_wait(&sh->wempty[s], ((cc / TC_NRING) & 1) ^ 1, sh->abort_flag, code);
    tc::mbar_arrive_expect_tx(&sh->wfull[s], bytes);
    tc::bulk_g2s(smem + TC_OFF_RING + s * TC_SLOT, wblob + src_off, bytes, &sh->wfull[s]);
    ++cc;
}
// table-driven variant (MMA self test)
__device__ __forceinline__ void tc_load_step_dyn(int st, uint32_t& cc, unsigned char* smem, const unsigned char* wblob) {
    const TcStep S = c_prog.steps[st];
#pragma unroll 1
    for (int c = 0; c < S.nchunks; ++c) {
        const TcChunk ch = c_prog.chunks[S.chunk0 + c];
        tc_load_chunk(ch.src_off, ch.bytes, cc, smem, wblob, 300 + st);
    }
}

// (Measured and dropped: a table-driven issuer - a ~100-instruction loop over a constant op table, records fetched one op
// ahead, K steps issued as one burst - instead of the ~55 KB of unrolled issue code below: 97.7 vs 92.3 ms per view.  The
// issue latency between "operands ready" and the last MMA of a step is on every tile's critical path, and immediates beat
// table lookups there even though the unrolled code is cold.)
// MMA issuer warp of tile group TG (whole warp, converged).  TG is a template parameter and every counter derives from
// kernel parameters and block indices, so that the compiler keeps the whole address arithmetic on the uniform datapath.
template <bool SPLIT>
__device__ __noinline__ void tc_issuer_warp(int tg_in, TcShared* sh, int V, int n_pairs) {
    const int tg = __shfl_sync(0xffffffffu, tg_in, 0);          // warp-uniform: the tile's offsets stay on the uniform datapath
    const int TG = tg;
    const uint32_t act_u32 = tc::smem_u32(tc_smem_raw) + TG * (TC_NACT * TC_SLOT), ring_u32 = tc::smem_u32(tc_smem_raw) + TC_OFF_RING;
    const uint32_t tmem = __shfl_sync(0xffffffffu, sh->tmem_base, 0) + TG * TC_TMEM_TILE;
    const bool lead = tc_elect();
    constexpr uint32_t cc_gm = kProg.cc_gm, cc_q = kProg.cc_q, cc_t = kProg.cc_t, cc_i = kProg.cc_i;
    uint32_t n = 0, cc = 0;
#define ISTEP(ST, COMMIT) tc_issuer_step<ST, COMMIT, SPLIT>(n, cc, sh, act_u32, ring_u32, tmem, tg, lead)
#pragma unroll 1
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        if (*reinterpret_cast<volatile int*>(sh->abort_flag)) break;
#pragma unroll 1
        for (int v = 0; v < V; ++v, cc += cc_gm) {
            ISTEP(ST_G1, 0); ISTEP(ST_G3, 0);
#if !TC_MERGE_G4
            ISTEP(ST_G4, 0);
#endif
            ISTEP(ST_M0, -1); ISTEP(ST_P0, 1); ISTEP(ST_P1, 2); ISTEP(ST_P2, 3); ISTEP(ST_P3, 1); ISTEP(ST_P4, 2); ISTEP(ST_P5, 0);
            ISTEP(ST_M1, 0); ISTEP(ST_M2, 0); ISTEP(ST_M3, 0);
        }
        ISTEP(ST_Q1, 0); ISTEP(ST_Q2, 0); ISTEP(ST_Q3, 0);
        cc += cc_q;
#pragma unroll 1
        for (int v = 0; v < V; ++v, cc += cc_t) { ISTEP(ST_T1, 0); ISTEP(ST_T2, 0); ISTEP(ST_T3, 0); ISTEP(ST_T4, 0); }
        // the rendering head runs once per tile for all views
        ISTEP(ST_I1, 0); ISTEP(ST_I2, 0); ISTEP(ST_I3, 0); ISTEP(ST_I4, 0); ISTEP(ST_I5, 0); ISTEP(ST_I6, 0); ISTEP(ST_I7, 0); ISTEP(ST_I8, 0);
#if !TC_I9_REGS
        ISTEP(ST_I9, 0);
#endif
        cc += cc_i;
    }
#undef ISTEP
}

// SPLIT = false: the bf16-MLP path, two tiles per CTA.  SPLIT = true: the split-precision ("fp32") path, one tile per CTA, the
// other tile group's warps idle (its operand slots hold the lo images).
#if TC_TAB_PARAM
template <bool SPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1) k_mlp_tc(const __grid_constant__ TcArgs A, const __grid_constant__ TcTables TAB) {
#else
template <bool SPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1) k_mlp_tc(TcArgs A) {
#endif
    unsigned char* smem = tc_smem_raw;
    TcShared* sh = reinterpret_cast<TcShared*>(smem + TC_OFF_CTRL);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int V = A.V;
    const int n_tiles = (A.n_chunk + TC_ROWS - 1) / TC_ROWS;
    constexpr int TPC = SPLIT ? 1 : TC_TILES;                  // tiles in flight per CTA
    constexpr int AUXB = SPLIT ? TC_AUX_BYTES_SPLIT : TC_AUX_BYTES;
    constexpr int NIMG = SPLIT ? 2 * TC_REC_IMAGES : TC_REC_IMAGES;    // operand images per (tile, view): hi [+ lo]
    const int n_pairs = (n_tiles + TPC - 1) / TPC;
#if TC_TAB_PARAM
    tc_setup(smem, sh, nullptr, TPC);
#else
    tc_setup(smem, sh, A.tab, TPC);
#endif

    // Register re-balancing (setmaxnreg, per warpgroup): 20 warps are launched at 96 registers so that the register
    // file of every SM sub-partition (5 warps x 32 x 96 <= 16 K) holds them; the last warpgroup (producer + 3 parked
    // warps) shrinks to 24 and the 16 tile warps grow to 112, which is what the epilogues need to stay out of local memory.
    if (warp >= TC_TILES * 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_PROD));
        // ===================================================== weight producer
        if (warp == TC_TILES * 8 + 1 || (!SPLIT && warp == TC_TILES * 8 + 2)) tc_issuer_warp<SPLIT>(warp - (TC_TILES * 8 + 1), sh, V, n_pairs);
        else if (warp == TC_TILES * 8 && lane == 0) {
            uint32_t cc = 0;
#pragma unroll 1
            for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                if (*reinterpret_cast<volatile int*>(sh->abort_flag)) break;        // a bounded wait gave up: drain
                // phases: 0 = geometry + MLP chunks (once per view), 1 = density head, 2 = texture fusion (per view), 3 = head
#pragma unroll 1
                for (int ph = 0; ph < 4; ++ph) {
                    const int first = ph == 0 ? 0 : ph == 1 ? kLoads.n_gm : ph == 2 ? kLoads.n_gm + kLoads.n_q : kLoads.n_gm + kLoads.n_q + kLoads.n_t;
                    const int cnt = ph == 0 ? kLoads.n_gm : ph == 1 ? kLoads.n_q : ph == 2 ? kLoads.n_t : kLoads.n_i;
                    const int reps = (ph == 0 || ph == 2) ? V : 1;
#pragma unroll 1
                    for (int rep = 0; rep < reps; ++rep)
#pragma unroll 1
                        for (int i = first; i < first + cnt; ++i) {
                            tc_load_chunk(c_loads.src_off[i], c_loads.bytes[i], cc, smem, A.wblob, 300 + ph);
                            if (SPLIT) tc_load_chunk(c_loads.src_off[i], c_loads.bytes[i], cc, smem, A.wblob + A.wblob_lo_off, 300 + ph);
                        }
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REGS_EPI));
        // ===================================================== tile threads
        TcTile t;
#if TC_TAB_PARAM
        tc_tile_init(t, smem, sh, lane, &TAB);
#else
        tc_tile_init(t, smem, sh, lane);
#endif
        const int row = t.row, h = t.half, tg = t.tg;
        const bool leader = tid == tg * TC_EPI_THREADS;          // issues this tile's record loads
#pragma unroll 1
        for (int pair = blockIdx.x; pair < (SPLIT && tg != 0 ? 0 : n_pairs); pair += gridDim.x) {
            // an odd tile count leaves the last pair's second group without a tile: it re-runs the last tile (the ring
            // needs both consumers) and stores nothing
            const int tile_raw = pair * TPC + tg;
            const int tile = min(tile_raw, n_tiles - 1);
            const int isamp = tile_raw < n_tiles ? tile * TC_ROWS + row : A.n_chunk;
            const unsigned char* aux_row = A.aux + ((size_t)tile * TC_ROWS + row) * V * AUXB;
            float wsum = 0.0f;
            // Operand images of (tile, view) -> slots 0..3 / texture image -> slot 1, requested by the group's leader at the point of use,
            // when every MMA that reads the target slots has completed.  (Requesting them a step early does not pay: section 9 of DESIGN.md.)
            auto load_geo = [&](int tl, int v) {
                if (TC_ABLATE & 128) return;
                const unsigned char* rimg = A.rec + ((size_t)tl * V + v) * (NIMG * TC_SLOT);
                tc::mbar_arrive_expect_tx(&sh->rec_bar[tg], (SPLIT ? 8 : 4) * TC_SLOT);
#pragma unroll 1
                for (int s = 0; s < 4; ++s) {
                    tc::bulk_g2s(t.slot(s), rimg + s * TC_SLOT, TC_SLOT, &sh->rec_bar[tg]);
                    if (SPLIT) tc::bulk_g2s(t.slot(s) + TC_LO_OFF, rimg + (TC_REC_IMAGES + s) * TC_SLOT, TC_SLOT, &sh->rec_bar[tg]);
                }
            };
            auto load_tex = [&](int v) {
                if (TC_ABLATE & 128) return;
                const unsigned char* rimg = A.rec + ((size_t)tile * V + v) * (NIMG * TC_SLOT);
                tc::mbar_arrive_expect_tx(&sh->rec_bar[tg], (SPLIT ? 2 : 1) * TC_SLOT);
                tc::bulk_g2s(t.slot(1), rimg + 4 * TC_SLOT, TC_SLOT, &sh->rec_bar[tg]);
                if (SPLIT) tc::bulk_g2s(t.slot(1) + TC_LO_OFF, rimg + (TC_REC_IMAGES + 4) * TC_SLOT, TC_SLOT, &sh->rec_bar[tg]);
            };
            // =========================================================== per-view geometry branch
#pragma unroll 1
            for (int v = 0; v < V; ++v) {
                // all MMAs that read slots 0..3 have completed (last wait_acc)
                if (leader) load_geo(tile, v);
                const float4 a0 = *reinterpret_cast<const float4*>(aux_row + v * AUXB);
                const float pw = reinterpret_cast<const float*>(aux_row + v * AUXB)[7];
                TC_PROF(5000 + v);
                if (!(TC_ABLATE & 128)) tc_wait(&sh->rec_bar[tg], t.rec_phase, sh->abort_flag, 400);
                t.rec_phase ^= 1;
                TC_PROF(5100 + v);
                // ---- G1: attention layer 1 of both scales -> acc cols 0..15 (scale 64), 16..31 (scale 8)
                t.step(ST_G1);
                // ---- attention layer 2 (10 -> 3, twice) + sigmoid in registers (fp32 weights), gates applied in place
                {
                    float hid[16];
                    float g64[3], g8[3];
                    t.ld16(0, hid);
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const float* w = t.tb->at2 + j * 12;
                        float s = 0.0f;
#pragma unroll
                        for (int i = 0; i < 10; ++i) s = fmaf(w[i], fmaxf(hid[i], 0.0f), s);
                        g64[j] = tc_act<TA_SIGMOID>(s);
                    }
                    t.ld16(16, hid);
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const float* w = t.tb->at2 + 36 + j * 12;
                        float s = 0.0f;
#pragma unroll
                        for (int i = 0; i < 10; ++i) s = fmaf(w[i], fmaxf(hid[i], 0.0f), s);
                        g8[j] = tc_act<TA_SIGMOID>(s);
                    }
#pragma unroll
                    for (int s = 0; s < 3; ++s)
#pragma unroll
                        for (int c = 0; c < 4; ++c) t.template gate_chunk1<SPLIT>(s, 4 * h + c, g64[s]);
                    if (h == 0) t.template gate_chunk1<SPLIT>(3, 2, g8[0]);
                    else { t.template gate_chunk1<SPLIT>(3, 3, g8[1]); t.template gate_chunk1<SPLIT>(3, 4, g8[2]); }
                }
                // ---- G3: fused layer 1 (ReLU)
                t.step(ST_G3);
                EPI(TA_RELU, 32 * h, 2, NOBIAS, 4, 4 * h);
                if (h == 0) EPI(TA_RELU, 64, 1, NOBIAS, 3, 6);
#if !TC_MERGE_G4
                // ---- G4: fused layer 2 -> out64 (slot 0), out8 (slot 3 cols 48..63)
                t.step(ST_G4);
                EPI(TA_NONE, 32 * h, 2, NOBIAS, 0, 4 * h);
                if (h == 1) EPI(TA_NONE, 64, 1, NOBIAS, 3, 6);
#endif
                // (TC_MERGE_G4: fused layer 2 is linear, bias-free and feeds only Linear layers, so its matrix is folded into their
                //  weights when they are packed - tc_build - and MLP layer 0 / 2 read G3's ReLU outputs in slot 4 / slot 3 directly)
                // ---- MLP layer 0: out64 part + the positional encoding (42 keypoints x 8 columns = six operand blocks, src/spatial.py:59-117)
                // PE block s (keypoints 8 s .. 8 s + 7; block 5: two keypoints) -> operand slot `pslot`, this thread's keypoints 8 s + 2 j + h
                auto gen_pe = [&](int s, int pslot) {
                    const int nk = s < 5 ? 4 : 1;
#pragma unroll 1
                    for (int j = 0; j < nk; ++j) {
                        const int kp = 8 * s + 2 * j + h;
                        const float4 kc = reinterpret_cast<const float4*>(t.tb->kpt4)[v * NKPT + kp];
                        const float dx = a0.x - kc.x, dy = a0.y - kc.y, dz = a0.z - kc.z;
#if TC_ABLATE & 2
                        const float w = (dx * dx + dy * dy + dz * dz) * -72.1347520f;
                        float s1 = dz, c1 = dx;
#else
                        const float w = tc_ex2((dx * dx + dy * dy + dz * dz) * -72.1347520f);   // exp(-d^2 / (2 * 0.1^2))
                        float s1, c1;
                        __sincosf(3.14159274f * dz, &s1, &c1);
#endif
                        const float s2 = 2.0f * s1 * c1, c2 = 1.0f - 2.0f * s1 * s1;
                        const float s4 = 2.0f * s2 * c2, c4 = 1.0f - 2.0f * s2 * s2;
                        if constexpr (SPLIT) t.template put8<true>(pslot, 2 * j + h, dz * w, s1 * w, c1 * w, s2 * w, c2 * w, s4 * w, c4 * w, 0.0f);
                        else t.st_chunk(pslot, 2 * j + h, make_uint4(tc_pack_scaled(dz, s1, w), tc_pack_scaled(c1, s2, w),
                                                                    tc_pack_scaled(c2, s4, w), tc::pack_bf16(c4 * w, 0.0f)));
                    }
                };
                // seven publishes: out64, then the six blocks through a 3-slot ring
                t.issue(ST_M0, -1);
#pragma unroll 1
                for (int s = 0; s < 6; ++s) {
                    const int ps = s % 3;
                    const int pslot = ps == 0 ? 1 : ps == 1 ? 2 : TC_PE_SLOT2;
                    if (s >= 3) {
                        tc_wait(&sh->pfree[tg][ps], (t.pfree_bits >> ps) & 1u, sh->abort_flag, 500 + s);
                        t.pfree_bits ^= 1u << ps;
                    }
                    gen_pe(s, pslot);
                    // P0..P4 commit to their ring-slot barrier, P5 completes the accumulator
                    t.issue(ST_P0 + s, s < 5 ? 1 + ps : 0);
                }
                t.wait_acc(ST_P5);
                // drain the ring-slot barriers committed by P3, P4 (slots 0, 1): already complete (MMAs complete in order)
#pragma unroll 1
                for (int ps = 0; ps < 2; ++ps) {
                    tc_wait(&sh->pfree[tg][ps], (t.pfree_bits >> ps) & 1u, sh->abort_flag, 510 + ps);
                    t.pfree_bits ^= 1u << ps;
                }
#if TC_ROLL_MLP
                // h0 -> slots 1, 2; h1 -> slots 4, 0; h2 -> slots 1, 2: one copy of the 64-column Softplus epilogue
#pragma unroll 1
                for (int l = 0; l < 3; ++l) {
                    EPI(TA_SOFTPLUS, 64 * h, 4, BIASP(L_MLP0 + l) + 64 * h, l == 1 ? (h ? 0 : 4) : 1 + h, 0);
                    t.step(ST_M1 + l);
                }
#else
                EPI(TA_SOFTPLUS, 64 * h, 4, BIASP(L_MLP0) + 64 * h, 1 + h, 0);        // h0 -> slots 1, 2
                t.step(ST_M1);
                EPI(TA_SOFTPLUS, 64 * h, 4, BIASP(L_MLP1) + 64 * h, h ? 0 : 4, 0);   // h1 -> slots 4, 0
                t.step(ST_M2);
                EPI(TA_SOFTPLUS, 64 * h, 4, BIASP(L_MLP2) + 64 * h, 1 + h, 0);        // h2 -> slots 1, 2
                t.step(ST_M3);
#endif
                // ---- weighted pooling sums over views in TMEM: S1 += w h3, S2 += w h3^2 (pool_ops, src/utils.py:854-880)
#pragma unroll 1
                for (int g = 0; g < 2; ++g) {
                    const int c0 = 32 * h + 16 * g;
                    float x[16], s1[16], s2[16];
                    t.ld16(c0, x);
                    if (v > 0) { t.ld16(TC_SREG + c0, s1); t.ld16(TC_SREG + 64 + c0, s2); }
                    const float* b3 = BIASP(L_MLP3) + c0;
#if TC_F32X2
                    const float2 pw2 = make_float2(pw, pw);
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const float2 hv = __fadd2_rn(make_float2(x[i], x[i + 1]), make_float2(b3[i], b3[i + 1]));
                        const float2 wh = __fmul2_rn(pw2, hv);
                        const float2 a1 = __fadd2_rn(v > 0 ? make_float2(s1[i], s1[i + 1]) : make_float2(0.f, 0.f), wh);
                        const float2 a2 = __ffma2_rn(wh, hv, v > 0 ? make_float2(s2[i], s2[i + 1]) : make_float2(0.f, 0.f));
                        s1[i] = a1.x; s1[i + 1] = a1.y; s2[i] = a2.x; s2[i + 1] = a2.y;
                    }
#else
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float hv = x[i] + b3[i];
                        const float wh = pw * hv;
                        s1[i] = (v > 0 ? s1[i] : 0.0f) + wh;
                        s2[i] = fmaf(wh, hv, v > 0 ? s2[i] : 0.0f);
                    }
#endif
                    t.st16(TC_SREG + c0, s1);
                    t.st16(TC_SREG + 64 + c0, s2);
                }
                tc::tmem_st_wait();
                wsum += pw;
            }
            // =========================================================== pooled latent, density head, compression
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {
                const int c0 = 32 * h + 16 * g;
                float m[16], q[16];
                t.ld16(TC_SREG + c0, m);
                t.ld16(TC_SREG + 64 + c0, q);
#pragma unroll
                for (int i = 0; i < 16; ++i) q[i] = q[i] - m[i] * m[i] * (2.0f - wsum);
                if (A.dbg_latent && isamp < A.n_chunk) {
                    float* o = A.dbg_latent + (size_t)(A.sample0 + isamp) * 128;
#pragma unroll
                    for (int i = 0; i < 16; ++i) { o[c0 + i] = m[i]; o[64 + c0 + i] = q[i]; }
                }
                t.template put8<SPLIT>(1, (c0 >> 3), m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7]);
                t.template put8<SPLIT>(1, (c0 >> 3) + 1, m[8], m[9], m[10], m[11], m[12], m[13], m[14], m[15]);
                t.template put8<SPLIT>(2, (c0 >> 3), q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7]);
                t.template put8<SPLIT>(2, (c0 >> 3) + 1, q[8], q[9], q[10], q[11], q[12], q[13], q[14], q[15]);
            }
            t.step(ST_Q1);
            uint4 lat_a = make_uint4(0, 0, 0, 0), lat_b = make_uint4(0, 0, 0, 0);      // h=0: latent cols 0-15, h=1: cols 16-23
            uint4 lat_al = make_uint4(0, 0, 0, 0), lat_bl = make_uint4(0, 0, 0, 0);    // their lo parts (split path)
            auto split8 = [](const float* x, uint4& hi, uint4& lo) {
                const float f[8] = {x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]};
                hi = pack8(f);
                float fh[8], fl[8];
                unpack8(hi, fh);
#pragma unroll
                for (int i = 0; i < 8; ++i) fl[i] = f[i] - fh[i];
                lo = pack8(fl);
            };
            EPI(TA_SOFTPLUS, 32 * h, 2, BIASP(L_POST0) + 32 * h, 4, 4 * h);
            if (h == 0) {
                float x[16];
                t.ld16(64, x);
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] += BIASP(L_COMPRESS)[i];
                split8(x, lat_a, lat_al);
                split8(x + 8, lat_b, lat_bl);
            } else {
                float x[8];
                t.ld8(80, x);
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] += BIASP(L_COMPRESS)[16 + i];
                split8(x, lat_a, lat_al);
            }
            t.step(ST_Q2);
            float o0, o1;
            EPI(TA_SOFTPLUS, 32 * h, 2, BIASP(L_POST1) + 32 * h, 0, 4 * h);
            t.step(ST_Q3);
            {
                float x[8];
                t.ld8(0, x);
                o0 = x[0] + BIASP(L_POST2)[0];
                o1 = x[1] + BIASP(L_POST2)[1];
            }
            // =========================================================== texture branch per view
#pragma unroll 1
            for (int v = 0; v < V; ++v) {
                if (leader) load_tex(v);
                // tail operand [lat24 | extras 8] in slot 2, ray difference (4) in slot 3 cols 0..15
                if (h == 0) {
                    const float4 a0 = *reinterpret_cast<const float4*>(aux_row + v * AUXB);
                    const float4 a1 = *reinterpret_cast<const float4*>(aux_row + v * AUXB + 16);
                    t.st_chunk(2, 0, lat_a);
                    t.st_chunk(2, 1, lat_b);
                    if (SPLIT) { t.st_chunk_lo(2, 0, lat_al); t.st_chunk_lo(2, 1, lat_bl); }
                    t.template put8<SPLIT>(3, 0, a0.w, a1.x, a1.y, a1.z, 0.0f, 0.0f, 0.0f, 0.0f);
                    t.template zero8<SPLIT>(3, 1);
                } else {
                    t.st_chunk(2, 2, lat_a);
                    if constexpr (SPLIT) {        // the 8 extras [a g, a b, b r, b g, b b, qvis, vn, vt] travel as fp32 in the split side record
                        t.st_chunk_lo(2, 2, lat_al);
                        const float4 e0 = *reinterpret_cast<const float4*>(aux_row + v * AUXB + 48);
                        const float4 e1 = *reinterpret_cast<const float4*>(aux_row + v * AUXB + 64);
                        t.template put8<true>(2, 3, e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w);
                    } else {
                        t.st_chunk(2, 3, *reinterpret_cast<const uint4*>(aux_row + v * AUXB + 48));
                    }
                }
                TC_PROF(5200 + v);
                if (!(TC_ABLATE & 128)) tc_wait(&sh->rec_bar[tg], t.rec_phase, sh->abort_flag, 401);
                t.rec_phase ^= 1;
                TC_PROF(5300 + v);
                // ---- T1: attention layer 1 (ReLU) + ray encoder layer 1 (ELU)
                t.step(ST_T1);
                if (h == 0) {
                    EPI(TA_RELU, 0, 3, NOBIAS, 4, 0);
                    EPI(TA_ELU, 96, 1, BIASP(L_RAY0), 3, 2);
                } else {
                    EPI(TA_RELU, 48, 1, NOBIAS, 4, 6);
                    EPI(TA_RELU, 64, 2, NOBIAS, 0, 0);
                }
                // ---- T2: attention layer 2 -> 6 sigmoid gates applied in place; ray encoder layer 2 -> dir40 to TMEM
                t.step(ST_T2);
                {
                    float gt[8];
                    t.ld8(0, gt);
#pragma unroll
                    for (int i = 0; i < 6; ++i) gt[i] = tc_act<TA_SIGMOID>(gt[i]);
                    // slot 1 chunks: 0 -> g0, 1 -> g1, 2 -> g2, 3,4 -> g3, 5,6 -> g4, 7 -> [g3,g3,g4,g4,g0,g0,g0,g1]
                    // slot 2 chunks: 0..2 -> g5, 3 -> [g1,g1,g2,g2,g2,1,1,1]      (kTexMap1 / kTexMap2)
                    if (h == 0) {
                        t.template gate_chunk1<SPLIT>(1, 0, gt[0]); t.template gate_chunk1<SPLIT>(1, 1, gt[1]); t.template gate_chunk1<SPLIT>(1, 2, gt[2]); t.template gate_chunk1<SPLIT>(1, 3, gt[3]);
                        t.template gate_chunk1<SPLIT>(2, 0, gt[5]); t.template gate_chunk1<SPLIT>(2, 1, gt[5]);
                    } else {
                        t.template gate_chunk1<SPLIT>(1, 4, gt[3]); t.template gate_chunk1<SPLIT>(1, 5, gt[4]); t.template gate_chunk1<SPLIT>(1, 6, gt[4]);
                        const float g7[8] = {gt[3], gt[3], gt[4], gt[4], gt[0], gt[0], gt[0], gt[1]};
                        t.template gate_chunk<SPLIT>(1, 7, g7);
                        t.template gate_chunk1<SPLIT>(2, 2, gt[5]);
                        const float g3[8] = {gt[1], gt[1], gt[2], gt[2], gt[2], 1.0f, 1.0f, 1.0f};
                        t.template gate_chunk<SPLIT>(2, 3, g3);
                    }
                    const int fcol = TC_SREG + 40 * v;
                    if (h == 0) {
                        float d[16], e[8];
                        t.ld16(16, d);
                        t.ld8(32, e);
#pragma unroll
                        for (int i = 0; i < 16; ++i) d[i] = tc_act<TA_ELU>(d[i] + BIASP(L_RAY1)[i]);
#pragma unroll
                        for (int i = 0; i < 8; ++i) e[i] = tc_act<TA_ELU>(e[i] + BIASP(L_RAY1)[16 + i]);
                        t.st16(fcol, d);
                        t.st8(fcol + 16, e);
                    } else {
                        float d[16];
                        t.ld16(40, d);
#pragma unroll
                        for (int i = 0; i < 16; ++i) d[i] = tc_act<TA_ELU>(d[i] + BIASP(L_RAY1)[24 + i]);
                        t.st16(fcol + 24, d);
                    }
                    tc::tmem_st_wait();
                }
                // ---- T3: fused layer 1 (ReLU)
                t.step(ST_T3);
                if (h == 0) EPI(TA_RELU, 0, 3, NOBIAS, 4, 0);
                else {
                    EPI(TA_RELU, 48, 1, NOBIAS, 4, 6);
                    EPI(TA_RELU, 64, 2, NOBIAS, 0, 0);
                }
                // ---- T4: fused layer 2 -> rgb_feat (40); source colour = channels 0..2; f = rgb_feat + dir
                t.step(ST_T4);
                {
                    const int fcol = TC_SREG + 40 * v;
                    if (h == 0) {
                        float x[16], y[8], d[16], e[8];
                        t.ld16(0, x);
                        t.ld8(16, y);
                        t.ld16(fcol, d);
                        t.ld8(fcol + 16, e);
                        const float src[8] = {x[0], x[1], x[2], 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int i = 0; i < 16; ++i) d[i] += x[i];
#pragma unroll
                        for (int i = 0; i < 8; ++i) e[i] += y[i];
                        t.st16(fcol, d);
                        t.st8(fcol + 16, e);
                        t.st8(TC_SRC + 4 * v, src);        // 4 columns per view; the 8-wide store only overlaps later views
                    } else {
                        float x[16], d[16];
                        t.ld16(24, x);
                        t.ld16(fcol + 24, d);
#pragma unroll
                        for (int i = 0; i < 16; ++i) d[i] += x[i];
                        t.st16(fcol + 24, d);
                    }
                    tc::tmem_st_wait();
                }
            }
            // =========================================================== IBRRenderingHead over the view axis
            float wt[TC_MAXV], maskv = 0.f, sv[TC_MAXV];
            {
                float e[TC_MAXV], emin = 3.0e38f, sum = 0.f;
#pragma unroll
                for (int v = 0; v < TC_MAXV; ++v) {
                    e[v] = 0.f; wt[v] = 0.f; sv[v] = -1e4f;
                    if (v < V) {
                        const float4 a1 = *reinterpret_cast<const float4*>(aux_row + v * AUXB + 16);
                        const float4 a2 = *reinterpret_cast<const float4*>(aux_row + v * AUXB + 32);
                        maskv = a2.x;
                        e[v] = __expf(t.tb->ani_al_abs * (a1.z - 1.0f));
                        emin = fminf(emin, e[v]);
                    }
                }
#pragma unroll
                for (int v = 0; v < TC_MAXV; ++v) if (v < V) { e[v] = (e[v] - emin) * maskv; sum += e[v]; }
#pragma unroll
                for (int v = 0; v < TC_MAXV; ++v) if (v < V) wt[v] = e[v] / (sum + 1e-8f);
            }
            // source colours of the views leave TMEM now: the batched steps below use accumulator columns 0..191
            float srcc[3 * TC_MAXV];
            {
                float s16[16];
                t.ld16(TC_SRC, s16);
#pragma unroll
                for (int v = 0; v < TC_MAXV; ++v)
#pragma unroll
                    for (int c = 0; c < 3; ++c) srcc[3 * v + c] = s16[4 * v + c];
            }
            // The nine steps of the head run ONCE per tile for all views: view v has its own operand slot (2 + v), its own
            // accumulator columns (stride 64 / 32 / 48 / 16 per step) and its residual x in TMEM columns 160 + 32 v; the
            // weights are shared.  Layout of slot 2 + v over the steps: I1 in [var 24..39 | f_v | pad], I2 in ELU(base0) 64,
            // I3 / I5 in cols 0..31, I4 / I6 in cols 32..63, I7 in cols 0..47, I8 in cols 48..63, I9 in cols 0..15.
            {   // fused_mean_variance (src/utils.py:153-157): channels 0..23 (h=0) / 24..39 (h=1), 8 at a time;
                // each thread reads back only the f columns it stored itself
                const int g0 = h ? 3 : 0, g1 = h ? 5 : 3;
#pragma unroll 1
                for (int g = g0; g < g1; ++g) {
                    float f[TC_MAXV][8], mean[8], var[8];
#pragma unroll
                    for (int v = 0; v < TC_MAXV; ++v) {
                        if (v < V) t.ld8(TC_SREG + 40 * v + 8 * g, f[v]);
                        else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) f[v][i] = 0.f;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float mu = 0.f;
#pragma unroll
                        for (int v = 0; v < TC_MAXV; ++v) mu += f[v][i] * wt[v];
                        float va = 0.f;
#pragma unroll
                        for (int v = 0; v < TC_MAXV; ++v) { const float d = f[v][i] - mu; va += wt[v] * d * d; }
                        mean[i] = mu; var[i] = va;
                    }
                    t.template put8<SPLIT>(1, g, mean);                               // mean -> slot 1 cols 0..39
                    if (g < 3) t.template put8<SPLIT>(1, 5 + g, var);                 // var 0..23 -> slot 1 cols 40..63
#pragma unroll
                    for (int v = 0; v < TC_MAXV; ++v) {
                        if (v < V) {
                            if (g >= 3) t.template put8<SPLIT>(2 + v, g - 3, var);    // var 24..39 -> slot 2+v cols 0..15
                            t.template put8<SPLIT>(2 + v, 2 + g, f[v]);               // f_v -> slot 2+v cols 16..55
                        }
                    }
                }
                if (h == 1) {
#pragma unroll
                    for (int v = 0; v < TC_MAXV; ++v) if (v < V) t.template zero8<SPLIT>(2 + v, 7);
                }
            }
            // all reads of the f columns precede the accumulator writes of I1 (ordered by the step's publish)
            t.step(ST_I1);
            TC_VLOOP
            for (int v = 0; v < TC_MAXV; ++v)
                if (v < V) tc_epi_store<TA_ELU, 2, SPLIT>(t, 64 * v + 32 * h, BIASP(L_BASE0) + 32 * h, t.slot(2 + v), 4 * h);
            t.step(ST_I2);
            TC_VLOOP
            for (int v = 0; v < TC_MAXV; ++v) {
                if (v < V) {
                    float x[16], y[16];
                    const float wv = tc_sel3(wt, v);
                    t.ld16(32 * v + 16 * h, x);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        x[i] = tc_act<TA_ELU>(x[i] + BIASP(L_BASE1)[16 * h + i]);
                        y[i] = x[i] * wv;
                    }
                    t.st16(160 + 32 * v + 16 * h, x);
                    t.template put8<SPLIT>(2 + v, 2 * h, y[0], y[1], y[2], y[3], y[4], y[5], y[6], y[7]);
                    t.template put8<SPLIT>(2 + v, 2 * h + 1, y[8], y[9], y[10], y[11], y[12], y[13], y[14], y[15]);
                }
            }
            tc::tmem_st_wait();
            t.step(ST_I3);
            TC_VLOOP
            for (int v = 0; v < TC_MAXV; ++v)
                if (v < V) tc_epi_store<TA_ELU, 1, SPLIT>(t, 48 * v + 16 * h, BIASP(L_VIS1_0) + 16 * h, t.slot(2 + v), 4 + 2 * h);
            t.step(ST_I4);
            TC_VLOOP
            for (int v = 0; v < TC_MAXV; ++v) {
                if (v < V) {
                    float r[16], vv[8], x[16], y[16];
                    t.ld16(48 * v + 16 * h, r);
                    t.ld8(48 * v + 32, vv);
                    t.ld16(160 + 32 * v + 16 * h, x);
                    const float vis = tc_act<TA_SIGMOID>(tc_act<TA_ELU>(vv[0] + BIASP(L_VIS1_1)[32])) * maskv;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        x[i] += tc_act<TA_ELU>(r[i] + BIASP(L_VIS1_1)[16 * h + i]);
                        y[i] = x[i] * vis;
                    }
                    t.st16(160 + 32 * v + 16 * h, x);
                    t.template put8<SPLIT>(2 + v, 2 * h, y[0], y[1], y[2], y[3], y[4], y[5], y[6], y[7]);
                    t.template put8<SPLIT>(2 + v, 2 * h + 1, y[8], y[9], y[10], y[11], y[12], y[13], y[14], y[15]);
                }
            }
            tc::tmem_st_wait();
            t.step(ST_I5);
            TC_VLOOP
            for (int v = 0; v < TC_MAXV; ++v)
                if (v < V) tc_epi_store<TA_ELU, 1, SPLIT>(t, 48 * v + 16 * h, BIASP(L_VIS2_0) + 16 * h, t.slot(2 + v), 4 + 2 * h);
            t.step(ST_I6);
            TC_VLOOP
            for (int v = 0; v < TC_MAXV; ++v) {
                if (v < V) {
                    float vv[8], x[16];
                    t.ld8(48 * v, vv);
                    t.ld16(160 + 32 * v + 16 * h, x);
                    const float vis2 = tc_act<TA_SIGMOID>(vv[0] + BIASP(L_VIS2_1)[0]) * maskv;
                    // out_layer input [x 32 | vis | ray_diff 4] -> slot 2+v cols 0..47
                    t.template put8<SPLIT>(2 + v, 2 * h, x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
                    t.template put8<SPLIT>(2 + v, 2 * h + 1, x[8], x[9], x[10], x[11], x[12], x[13], x[14], x[15]);
                    if (h == 0) {
                        const float4 a0 = *reinterpret_cast<const float4*>(aux_row + v * AUXB);
                        const float4 a1 = *reinterpret_cast<const float4*>(aux_row + v * AUXB + 16);
                        t.template put8<SPLIT>(2 + v, 4, vis2, a0.w, a1.x, a1.y, a1.z, 0.0f, 0.0f, 0.0f);
                    } else {
                        t.template zero8<SPLIT>(2 + v, 5);
                    }
                }
            }
            t.step(ST_I7);
            TC_VLOOP
            for (int v = 0; v < TC_MAXV; ++v)            // 16 columns per view: views alternate between the two row partners
                if (v < V && (v & 1) == h) tc_epi_store<TA_ELU, 1, SPLIT>(t, 16 * v, BIASP(L_OUT0), t.slot(2 + v), 6);
            t.step(ST_I8);
#if TC_I9_REGS
            // out_layer's last two layers end here: s_v = w2 . ELU(acc_v + b1) + b2 in fp32 registers (8 values per view)
            if (h == 0) {
                TC_VLOOP
                for (int v = 0; v < TC_MAXV; ++v) {
                    if (v < V) {
                        float x[8];
                        t.ld8(16 * v, x);
                        float sacc = BIASP(L_OUT2)[0];
#pragma unroll
                        for (int i = 0; i < 8; ++i) sacc = fmaf(t.tb->out2w[i], tc_act<TA_ELU>(x[i] + BIASP(L_OUT1)[i]), sacc);
                        tc_set3(sv, v, (maskv == 0.0f) ? -1e4f : sacc);
                    }
                }
            }
#else
            TC_VLOOP
            for (int v = 0; v < TC_MAXV; ++v)
                if (v < V && (v & 1) == h) tc_epi_store<TA_ELU, 1, SPLIT>(t, 16 * v, BIASP(L_OUT1), t.slot(2 + v), 0);
            t.step(ST_I9);
#endif
#if !TC_I9_REGS
            if (h == 0) {
                TC_VLOOP
                for (int v = 0; v < TC_MAXV; ++v) {
                    if (v < V) {
                        float vv[8];
                        t.ld8(16 * v, vv);
                        tc_set3(sv, v, (maskv == 0.0f) ? -1e4f : (vv[0] + BIASP(L_OUT2)[0]));
                    }
                }
            }
#endif
            // =========================================================== softmax blend + eval_func (src/model.py:1634-1635, 1140-1160)
            if (h == 0) {
                float smax = -3.0e38f;
#pragma unroll
                for (int v = 0; v < TC_MAXV; ++v) if (v < V) smax = fmaxf(smax, sv[v]);
                float den = 0.f, rgb[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int v = 0; v < TC_MAXV; ++v) {
                    if (v < V) {
                        const float e = __expf(sv[v] - smax);
                        den += e;
#pragma unroll
                        for (int c = 0; c < 3; ++c) rgb[c] += srcc[3 * v + c] * e;
                    }
                }
                const float inv = 1.0f / den;
                if (isamp < A.n_chunk) {
                    const size_t n = (size_t)(A.sample0 + isamp);
                    if (A.raw_out) {
                        float* o = A.raw_out + n * 5;
                        o[0] = o0; o[1] = o1; o[2] = rgb[0] * inv; o[3] = rgb[1] * inv; o[4] = rgb[2] * inv;
                    }
                    if (A.rgba) {
                        float* o = A.rgba + n * 5;
                        o[0] = maskv * fmaxf(o1, 0.0f);
                        o[1] = maskv * o0 + (1.0f - maskv) * 0.001f;
                        o[2] = rgb[0] * inv; o[3] = rgb[1] * inv; o[4] = rgb[2] * inv;
                    }
                }
            }
            TC_PROF(6000);
            // all TMEM reads of this tile precede the next tile's MMAs (ordered by the next step barrier)
            // leave the loop as a group once a bounded wait has given up (a named barrier has no time-out: the decision
            // must be the same for all 256 threads, so the leader takes it between two group barriers)
            tc::named_bar_sync(1 + tg, TC_EPI_THREADS);
            if (leader) sh->stop[tg] = *reinterpret_cast<volatile int*>(sh->abort_flag);
            tc::named_bar_sync(1 + tg, TC_EPI_THREADS);
            if (*reinterpret_cast<volatile int*>(&sh->stop[tg])) break;
        }
    }
    tc_teardown(sh, A.err);
}
// ================================================================================================ self test
// D (128 x Npad, fp32) = bf16(A (128 x K)) * bf16(W (N x K))^T through the same slots / ring / issue / TMEM path as
// k_mlp_tc, K <= 256 (multiple of 16), N <= 128.  c_tc holds a one-step table (index 0) built by tc_build_single.
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_selftest(const TcTables* tab, const unsigned char* wblob, const float* A_in, int K,
                                                               int n_pad, float* D_out, int* err) {
    unsigned char* smem = tc_smem_raw;
    TcShared* sh = reinterpret_cast<TcShared*>(smem + TC_OFF_CTRL);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    tc_setup(smem, sh, tab, 1);                  // only tile group 0 consumes the ring
    if (warp == TC_TILES * 8) {
        if (lane == 0) {
            uint32_t cc = 0;
            tc_load_step_dyn(0, cc, smem, wblob);
        }
    } else if (warp == TC_TILES * 8 + 1) {       // MMA issuer of tile group 0: one step
        const uint32_t tmem = __shfl_sync(0xffffffffu, sh->tmem_base, 0);
        tc::mbar_wait(&sh->ready[0][0], 0, sh->abort_flag, 600);
        tc::tcgen05_fence_after();
        tc_issue_ops_dyn(0, 0, sh, tc::smem_u32(tc_smem_raw), tc::smem_u32(tc_smem_raw) + TC_OFF_RING, tmem, 0);
    } else if (warp < 8) {
        TcTile t;
        tc_tile_init(t, smem, sh, lane);
        // operand: columns [64 s, 64 s + 64) of A -> slot s; this thread converts chunks 4 h .. 4 h + 3 of its row
        for (int s = 0; s * 64 < K; ++s)
            for (int c = 4 * t.half; c < 4 * t.half + 4; ++c) {
                float f[8];
                for (int i = 0; i < 8; ++i) {
                    const int k = 64 * s + 8 * c + i;
                    f[i] = k < K ? A_in[(size_t)t.row * K + k] : 0.0f;
                }
                t.st_chunk(s, c, pack8(f));
            }
        t.step(0);
        for (int g = t.half; g * 16 < n_pad; g += 2) {       // 16-column groups alternate between the two row partners
            const int c0 = 16 * g;
            float v[16];
            t.ld16(c0, v);
            for (int i = 0; i < 16; ++i) D_out[(size_t)t.row * n_pad + c0 + i] = v[i];
        }
    }
    tc_teardown(sh, err);
}

// One-step table for the self test: K columns over slots 0.., identity K order.
static void tc_build_single(const float* W, int N, int K, TcProg& T, std::vector<uint16_t>& blob) {
    memset(&T, 0, sizeof(T));
    blob.clear();
    const int n_pad = (N + 15) & ~15;
    int n_ops = 0, n_chunks = 0;
    uint32_t cur = 0;
    for (int k0 = 0; k0 < K; k0 += 64) {
        const int ncols = std::min(64, K - k0);
        const uint32_t bytes = (uint32_t)n_pad * 128;
        if (n_chunks == 0 || cur + bytes > TC_SLOT) {
            if (n_ops) T.ops[n_ops - 1].last_in_chunk = 1;
            T.chunks[n_chunks].src_off = (uint32_t)(blob.size() * 2);
            T.chunks[n_chunks].bytes = 0;
            ++n_chunks;
            cur = 0;
        }
        TcOp& d = T.ops[n_ops++];
        d.a_off = (uint32_t)((k0 / 64) * TC_SLOT);
        d.b_off = cur;
        d.idesc = tc::umma_idesc_bf16(128, n_pad);
        d.d_col = 0;
        d.nk = (uint8_t)(ncols / 16);
        d.accum = k0 > 0;
        d.chunk_rel = (uint8_t)(n_chunks - 1);
        d.last_in_chunk = 0;
        const size_t base = blob.size();
        blob.resize(base + bytes / 2, 0);
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < ncols; ++k) {
                const size_t byte = (size_t)(n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 3) ^ n) & 7) << 4) + (k & 7) * 2;
                blob[base + byte / 2] = f2bf_host(W[(size_t)n * K + k0 + k]);
            }
        cur += bytes;
        T.chunks[n_chunks - 1].bytes = cur;
    }
    T.ops[n_ops - 1].last_in_chunk = 1;
    T.steps[0].op0 = 0; T.steps[0].nops = (uint16_t)n_ops; T.steps[0].chunk0 = 0; T.steps[0].nchunks = (uint16_t)n_chunks;
}

// ================================================================================================ MMA pacing probe
// Developer measurement (vanerf_tc_mma_probe): one warp issues `reps` tcgen05.mma (M = 128, N = n, K = 16, operands =
// whatever the shared memory holds) round-robin over `n_acc` accumulators and reports, per CTA, the cycles spent
// issuing and the cycles until the last MMA has completed.  mode bit 0: a second warp issues the same stream into the
// other TMEM half concurrently (two tiles per SM).
__global__ void __launch_bounds__(128, 1) k_tc_mma_probe(int n, int reps, int n_acc, int mode, long long* out) {
    unsigned char* smem = tc_smem_raw;
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_slot;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    for (int i = threadIdx.x; i < (6 * TC_SLOT) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { tc::mbar_init(&bar[0], 1); tc::mbar_init(&bar[1], 1); tc::mbar_fence_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    tc::fence_proxy_async();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    if (warp < 2 && (warp == 0 || (mode & 1))) {
        const uint32_t base = tc::smem_u32(smem) + warp * 2 * TC_SLOT;
        const uint32_t idesc = tc::umma_idesc_bf16(128, n);
        const bool lead = tc_elect();
        const long long t0 = clock64();
        const uint32_t a0 = tmem + warp * 256, a1 = a0 + (uint32_t)(1 % n_acc) * n, a2 = a0 + (uint32_t)(2 % n_acc) * n, a3 = a0 + (uint32_t)(3 % n_acc) * n;
        const uint64_t ad = tc_desc(base), bd = tc_desc(base + TC_SLOT);
        if (lead) {
#pragma unroll 1
            for (int r = 0; r < reps; r += 4) {          // four K steps of one 64-column operand slot per iteration
                tc::umma_bf16(a0, ad, bd, idesc, 1u);
                tc::umma_bf16(a1, ad + 2, bd + 2, idesc, 1u);
                tc::umma_bf16(a2, ad + 4, bd + 4, idesc, 1u);
                tc::umma_bf16(a3, ad + 6, bd + 6, idesc, 1u);
            }
            tc::umma_commit(&bar[warp]);
        }
        __syncwarp();
        const long long t1 = clock64();
        while (!tc::mbar_try_wait(&bar[warp], 0)) {}
        const long long t2 = clock64();
        if (lead) {
            out[(blockIdx.x * 2 + warp) * 2] = t1 - t0;
            out[(blockIdx.x * 2 + warp) * 2 + 1] = t2 - t0;
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}
