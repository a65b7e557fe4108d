/* tetris_piclim.h -- C ABI of the B200-native Tetris-piclim rollout hot path (libtetris_piclim_sm100.so).
 *
 * Drop-in boundary for the hot path of the reference's game/tetris.py (all file:line citations are
 * relative to the reference tree): prescribed reset (:438-449), Tetris.move (:354-422) with
 * calculate_drop(_deltas) (:424-433) and get_tetromino (:60-61), Tetris.get_state (:435-436), plus the
 * afterstate enumeration + features and the counter-based 7-bag piece RNG the north star adds
 * (RandomPieceGenerator contract, :64-108).  Plain pointers and sizes only; no torch types.
 *
 * Two groups of entry points:
 *   tpl_*        device-pointer API: every array argument is a DEVICE pointer, work is queued
 *                asynchronously on `stream` (a cudaStream_t passed as void*); the library never
 *                allocates, frees or synchronises.  Stream order is what a caller sees; underneath, the
 *                persistent kernels (tpl_step, tpl_afterstates*, tpl_step_observe*) are launched with
 *                programmatic dependent launch, i.e. their set-up may overlap the tail of the previous
 *                kernel in the stream while everything that touches data waits for it to complete
 *                (TPL_NO_PDL=1 in the environment, read once per process, launches them plainly).
 *   tpl_env_*    host-buffer API: an opaque handle owns the device state, pinned staging and a
 *                stream; every array argument is a HOST pointer and the call returns when the
 *                outputs are valid.  This is what a ctypes binding in the reference would call
 *                (see INTEGRATION.md).
 *
 * All functions return 0 on success, a positive cudaError_t value for CUDA failures, or a negative
 * TPL_E* code for argument errors; tpl_last_error() returns a thread-local description.
 *
 * ---- data formats --------------------------------------------------------------------------------
 * Canonical (boundary) board: 20 x uint16 bitrows, bit c = column c, row 0 = top (reference board[0]),
 *   full row = 0x3FF.  Pieces: one byte each, 0=I 1=L 2=J 3=T 4=S 5=Z 6=O (:8-16, :23-57).
 * Env record (device, 64 bytes = 4 x 16-byte chunks), as 16 little-endian uint32 words:
 *   w[0..9]   board as 10 bit-columns: bit b of w[c] = cell (row 19-b, column c); bit 0 = floor row
 *   w[10..13] piece queue, 3 bits per piece, piece i at bits [3i, 3i+3) of the 128-bit value (<= 42 pieces)
 *   w[14]     lines_cleared (low 16 bits) | moves_used (high 16 bits)
 *   w[15]     state (byte 0: 0 running / 1 won / 2 lost  == reference None / True / False)
 *             | head (byte 1: pieces already popped) | npieces (byte 2)
 * State array ("planes"): chunk j of env i lives at ((uint4*)state)[j * plane_stride + i]; a warp reading
 *   chunk j of 32 consecutive envs reads 512 contiguous bytes.  Allocate 64 * plane_stride bytes.
 * Config pool: K records, array-of-structs: record k at ((uint4*)pool)[4*k .. 4*k+3].
 * Afterstate outputs, slot s = rot*10 + loc (rot 0..3, loc 0..9), slot-major so that stores coalesce:
 *   feats  uint8[40][n][4] = (rows cleared, holes, bumpiness, aggregate height)
 *   flags  uint8[40][n]    = TPL_FLAG_* bits
 *   feats_f32 float[40][n][4] (optional) = the same four numbers as floats (value-net input rows)
 * Distinct-placements form (tpl_*_distinct): the 40-slot grid aliases by rot % n_rot (:61) and min(loc, 10 - w) (:364), so
 *   only 9 (O), 17 (I, S, Z) or 34 (L, J, T) of the 40 slots differ -- 23.1 on average.  This form writes exactly those:
 *   rows  uint32[...]  one word per distinct placement, byte 0 = rows cleared | flags << 3, then holes, bumpiness, aggregate
 *                      height (the compact form's word; TPL_FLAG_ALIAS never set).  The placements of one env are contiguous
 *                      ("run"), ordered by rotation r < n_rot, then column c <= 10 - w.
 *   runs  uint32[n]    run descriptor of env i: TPL_RUN_OFFSET = word offset of its run in rows, TPL_RUN_PIECE = its current
 *                      piece, which fixes the run's length and the (rot, loc) of each placement (tpl_distinct_tables).
 *   Runs of a 32-env tile are adjacent; tiles are placed by an atomic counter, so their order in rows is unspecified (and
 *   may differ between runs of the same program); tiles start 16-byte aligned, gaps (< 4 words) hold zeros.
 */
#ifndef TETRIS_PICLIM_H_
#define TETRIS_PICLIM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPL_ABI_VERSION 2
#define TPL_MAX_PIECES 42
#define TPL_RECORD_BYTES 64
/* distinct-placements ("alias-free") afterstate form: at most 34 placements per env (L, J, T), see tpl_afterstates_distinct */
#define TPL_DISTINCT_MAX 34
#define TPL_DISTINCT_CAPACITY(n) (34 * (int64_t)(n) + 4 * (((int64_t)(n) + 31) / 32))   /* words the rows array must hold */
#define TPL_RUN_OFFSET(d) ((d) & 0x1FFFFFFFu)   /* run descriptor -> word offset of the env's first placement */
#define TPL_RUN_PIECE(d) ((d) >> 29)            /* run descriptor -> current piece 0..6, 7 = empty queue (run length 0) */

/* move / afterstate flags */
#define TPL_FLAG_TOPOUT 1    /* drop row < 0: piece consumed, board and moves_used unchanged, state lost (:372-374) */
#define TPL_FLAG_WIN 2       /* lines_cleared >= L on this clearing move (:415-417) */
#define TPL_FLAG_LOSE 4      /* moves_used >= M and no win on this move (:389-391, :420-422) */
#define TPL_FLAG_ALIAS 8     /* afterstates only: (rot, loc) wraps/clamps onto an earlier slot (:61, :364) */
#define TPL_FLAG_NOPIECE 16  /* piece queue empty: nothing done (the reference raises IndexError at :356) */

/* states */
#define TPL_RUNNING 0
#define TPL_WON 1
#define TPL_LOST 2

/* reset modes */
#define TPL_RESET_ALL 0      /* every env */
#define TPL_RESET_MASK 1     /* envs with mask[i] != 0 */
#define TPL_RESET_DONE 2     /* envs whose state != running or whose queue is empty (auto-reset) */

/* argument errors */
#define TPL_EINVAL (-1)
#define TPL_ERANGE (-2)
#define TPL_ENOMEM (-3)

int tpl_abi_version(void);
const char *tpl_last_error(void);

/* ---------------------------------------------------------------------------------------------------
 * device-pointer API
 * ------------------------------------------------------------------------------------------------- */

/* Install canonical (rows, pieces) data into records.  Replaces the reset-point hand-over at
 * game/tetris.py:447 (`self.board, self.pieces = self.queue.get()`) with ctor-fresh counters (:149-151).
 *   rows u16[n][20]; pieces u8[n][pieces_stride]; npieces u8[n] (each <= 42)
 *   lines/moves i32[n], st i8[n], head u8[n]: optional (NULL = 0) -- lets tests install mid-episode states
 *   aos != 0: write a config pool (record k at out[4k..4k+3]); aos == 0: write state planes. */
int tpl_pack(void *out, int64_t plane_stride, int aos, int n,
             const uint16_t *rows, const uint8_t *pieces, int pieces_stride, const uint8_t *npieces,
             const int32_t *lines, const int32_t *moves, const int8_t *st, const uint8_t *head, void *stream);

/* Tetris.get_state (:435-436) for a batch, plus the raw fields tests compare.  Every output is optional.
 *   rows u16[n][20]; cur/next u8[n] (255 = none); lines/moves i32[n] (cleared/used, NOT remaining);
 *   st i8[n]; head/npieces u8[n]; queue u8[n][42] = the full piece list including already-popped ones. */
int tpl_unpack(const void *state, int64_t plane_stride, int n, uint16_t *rows, uint8_t *cur, uint8_t *next,
               int32_t *lines, int32_t *moves, int8_t *st, uint8_t *head, uint8_t *npieces, uint8_t *queue,
               void *stream);

/* Tetris.reset + load_warm_reset (:438-449) for the prescribed-config half: copy pool records into envs and
 * zero lines/moves/state/head.  idx i32[n] picks the config per env; idx == NULL draws
 * k = mulhi(philox(seed; env_base+i, episode[i], CONFIG).w0, K).  mode: TPL_RESET_*.  episode u32[n] (optional):
 * incremented *before* the draw for every env reset in mode TPL_RESET_DONE, and in mode TPL_RESET_MASK when idx == NULL
 * (a reset that draws its config starts a new episode; with explicit idx the caller owns the numbering).  tstep u32[n]
 * (optional): the rollouts' per-episode action counter, zeroed for every env that is reset.  gen_count > 0 replaces the
 * pool's pieces by gen_count pieces of the counter-based 7-bag stream of (seed, env, episode). */
int tpl_reset_from_pool(void *state, int64_t plane_stride, int n, const void *pool, int K,
                        const int32_t *idx, const uint8_t *mask, int mode, uint32_t *episode, uint32_t *tstep,
                        uint64_t seed, uint64_t env_base, int gen_count, void *stream);

/* Tetris.move (:354-422) on every env: rot u8[n] (already reduced mod 4 by the caller; rot % n_rot is applied
 * here), loc u8[n] (clamped to 10 - width like :364).  Outputs (each optional): dlines i8[n] rows cleared,
 * flags u8[n] TPL_FLAG_*, st i8[n] state after the move, stats i64[8] += {episodes ended, wins, top-outs,
 * move-limit losses, lines, moves placed, steps, 0} (warp-reduced, one atomic per warp and counter). */
int tpl_step(void *state, int64_t plane_stride, int n, const uint8_t *rot, const uint8_t *loc,
             int8_t *dlines, uint8_t *flags, int8_t *st, long long *stats, int L, int M, void *stream);

/* Afterstate enumeration: slot (r, c) == clone(env).move(r, c) (composition of :354-422), with features on
 * the post-move board (the unchanged board when the move tops out).  Output forms (slot-major, see above):
 *   feats + flags            parity form, 200 B/env: feats byte 0 = rows cleared, flags separate
 *   feats only (flags NULL)  compact form, 160 B/env: feats byte 0 = rows cleared | flags << 3
 *   feats_f32 + flags        value-net form (float4 per slot); may be combined with feats
 * n <= 2^25 per call. */
int tpl_afterstates(const void *state, int64_t plane_stride, int n, uint8_t *feats, uint8_t *flags,
                    float *feats_f32, int L, int M, void *stream);

/* The fused hot-path step: tpl_step, then tpl_reset_from_pool(TPL_RESET_DONE) (skipped when pool == NULL), then
 * tpl_afterstates of the resulting state, in ONE kernel: each record is read once and written once.  Arguments as in
 * the three calls it replaces; results are identical to running them in sequence. */
int tpl_step_observe(void *state, int64_t plane_stride, int n, const uint8_t *rot, const uint8_t *loc,
                     int8_t *dlines, uint8_t *flags, int8_t *st, long long *stats,
                     const void *pool, int K, uint32_t *episode, uint32_t *tstep, uint64_t seed, uint64_t env_base, int gen_count,
                     uint8_t *feats, uint8_t *aflags, float *feats_f32, int L, int M, void *stream);

/* The distinct-placements form of tpl_afterstates / tpl_step_observe (see "data formats"): the same enumeration, but only the
 * placements that differ are written -- 92 instead of 160 bytes per env on average, and no duplicate rows for a value net.
 *   rows  u32[rows_capacity], 16-byte aligned, rows_capacity >= TPL_DISTINCT_CAPACITY(n) words
 *   runs  u32[n] run descriptors; run_base is added to the offsets they report (not to the addresses written), so several
 *         calls over sub-ranges of the envs can fill one array (tpl_env_step_observe_distinct does that)
 *   cursor2 u32[2] device counters, zero before the first call: the call appends at cursor2[phase] (afterwards the number of
 *         words used, gaps included) and clears cursor2[phase ^ 1]; alternate phase = 0, 1, 0, ... between consecutive calls
 *         on the same stream and no memset is ever needed.  n <= 2^23 per call. */
int tpl_afterstates_distinct(const void *state, int64_t plane_stride, int n, uint32_t *rows, int64_t rows_capacity,
                             uint32_t *runs, uint32_t run_base, uint32_t *cursor2, int phase, int L, int M, void *stream);
int tpl_step_observe_distinct(void *state, int64_t plane_stride, int n, const uint8_t *rot, const uint8_t *loc,
                              int8_t *dlines, uint8_t *flags, int8_t *st, long long *stats,
                              const void *pool, int K, uint32_t *episode, uint32_t *tstep, uint64_t seed, uint64_t env_base,
                              int gen_count, uint32_t *rows, int64_t rows_capacity, uint32_t *runs, uint32_t run_base,
                              uint32_t *cursor2, int phase, int L, int M, void *stream);
/* distinct placements -> the compact 40-slot form feats u8[40][n][4] (byte 0 = rows cleared | flags << 3, TPL_FLAG_ALIAS on the
 * slots that repeat an earlier one): what tpl_afterstates would have written.  Device pointers. */
int tpl_expand_distinct(const uint32_t *rows, const uint32_t *runs, int n, uint8_t *feats, void *stream);
/* Host tables of the numbering (any pointer may be NULL): count[7] = run length per piece; slot_of[7][34] = 40-slot index
 * rot * 10 + loc of placement j (255 past the run's end); canon_of[7][40] = placement index that slot (rot, loc) aliases. */
void tpl_distinct_tables(uint8_t *count, uint8_t *slot_of, uint8_t *canon_of);

/* Counter-based 7-bag piece sequences (contract of RandomPieceGenerator.get_random_sequence, :95-102):
 * out u8[n][count]; episode u32[n] or NULL (= episode0 for all). */
int tpl_gen_pieces(uint8_t *out, int n, int count, uint64_t seed, uint64_t env_base, const uint32_t *episode,
                   uint32_t episode0, void *stream);

/* Fused random-agent rollout (config 1's agent: rot ~ U{0..3}, loc ~ U{0..9} from the counter RNG), `steps`
 * moves per env with auto-reset from the pool, state kept in registers between moves.
 *   episode/tstep u32[n] in/out; stats i64[8] += {episodes, wins, top-outs, move-limit losses, lines,
 *   moves placed, steps, resets}. */
int tpl_rollout_random(void *state, int64_t plane_stride, int n, const void *pool, int K,
                       uint32_t *episode, uint32_t *tstep, long long *stats, int steps,
                       uint64_t seed, uint64_t env_base, int gen_count, int L, int M, void *stream);

/* Fused greedy rollout: every step enumerates the 40 afterstates in registers, scores each with the integer
 * linear value w[0]*dlines + w[1]*holes + w[2]*bumpiness + w[3]*agg_height (+ w[4] if it wins, + w[5] if it loses
 * or tops out), plays the arg-max slot (lowest slot on ties) and auto-resets.  weights6_host: HOST int32[6].
 * Same in/out arrays as above. */
int tpl_rollout_greedy(void *state, int64_t plane_stride, int n, const void *pool, int K,
                       uint32_t *episode, uint32_t *tstep, long long *stats, int steps, const int32_t *weights6_host,
                       uint64_t seed, uint64_t env_base, int gen_count, int L, int M, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * value net for ranking afterstates: the reference's layer shape model/model.py:9-20 as 4 -> 128 -> 128 -> 128
 * -> 128 -> 1, inference only, bf16 operands with fp32 accumulation on the tensor cores (tcgen05), weights and activations
 * on chip.  Training stays in PyTorch; these calls only serve the rollout's action selection.  Device pointers.
 * ------------------------------------------------------------------------------------------------- */
#define TPL_VALUE_BLOB_BYTES 119328
/* fp32 parameters in PyTorch layout (weight [out][in]: w1 [128][4], w2..w4 [128][128], w5 [1][128]; biases [128] / [1]) ->
 * the packed bf16 blob (TPL_VALUE_BLOB_BYTES, 16-byte aligned) tpl_value_rows reads.  scale4_host: HOST float[4], the factors the
 * four u8 features (rows cleared, holes, bumpiness, aggregate height) are multiplied by before the first layer. */
int tpl_value_pack(const float *w1, const float *b1, const float *w2, const float *b2, const float *w3, const float *b3,
                   const float *w4, const float *b4, const float *w5, const float *b5, const float *scale4_host, void *blob,
                   void *stream);
/* values[r] = V(rows[r]) for every r < min(*count, nrows_max) (count == NULL: r < nrows_max): rows are feature words in the
 * distinct-placements form (byte 0 & 7 = rows cleared, bytes 1-3 = holes, bumpiness, aggregate height); count is typically
 * cursor2[phase] of the tpl_*_distinct call that wrote them, read on the device -- no host synchronisation. */
int tpl_value_rows(const uint32_t *rows, const uint32_t *count, int64_t nrows_max, const void *blob, float *values, void *stream);
/* Per env, over its run of distinct placements: q = rows cleared + reward_win / reward_lose on a winning / losing (or
 * topping-out) placement + gamma * values; arg-max (lowest placement on ties), replaced with probability eps by a uniformly
 * random placement (counter RNG, stream 3, keyed by (seed, env_base + i, step)).  Outputs: rot / loc u8[n] (the action for
 * tpl_step*), chosen u32[n] (optional: the chosen placement's feature word), chosen_q f32[n] (optional: the arg-max q). */
int tpl_select_action(const uint32_t *rows, const uint32_t *runs, const float *values, int n, float gamma, float reward_win,
                      float reward_lose, float eps, uint64_t seed, uint64_t env_base, uint32_t step, uint8_t *rot, uint8_t *loc,
                      uint32_t *chosen, float *chosen_q, void *stream);

/* One transition per env from a distinct-form step into the ring buffers of a replay memory (slot = (pos + i) % capacity), in
 * the form a TD target consumes without further element-wise work:
 *   x u32 = the chosen placement's feature word (tpl_select_action) with the flag bits cleared, so that its four bytes ARE the
 *   features;  reward f32 = rows cleared + reward_win / reward_lose by the move's flags;  live f32 = 0 if the episode ended or
 *   nothing is left to place, else 1;  next_w u32[capacity][TPL_DISTINCT_MAX] = the new state's placements (flags cleared, zero
 *   padded);  next_r f32[..][34] = reward of each (-inf past the end of the run; next_r[0] = 0 for an empty run);
 *   next_g f32[..][34] = gamma where the placement does not end the episode, else 0 -- Q = next_r + next_g * V(next_w).
 * Device pointers; one launch instead of a few dozen framework ops per rollout step. */
int tpl_replay_push(const uint32_t *rows, const uint32_t *runs, const int8_t *dlines, const uint8_t *mflags, const int8_t *state,
                    const uint32_t *chosen, int n, int64_t pos, int64_t capacity, float gamma, float reward_win, float reward_lose,
                    uint32_t *x, float *reward, float *live, uint32_t *next_w, float *next_r, float *next_g, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * host-buffer API (handle owns device memory; every array is a HOST pointer)
 * ------------------------------------------------------------------------------------------------- */
typedef struct tpl_env tpl_env;

/* Tetris.__init__(L, M, ...) (:141-151) for n envs on CUDA device `device`. */
int tpl_env_create(tpl_env **out, int n, int L, int M, int device, uint64_t seed, uint64_t env_base);
void tpl_env_destroy(tpl_env *e);
/* the reference's L and M are public mutable attributes (:143-144): change them for subsequent calls */
int tpl_env_set_limits(tpl_env *e, int L, int M);
/* upload K prescribed reset points (the (board, pieces) tuples of :476-479) */
int tpl_env_set_pool(tpl_env *e, int K, const uint16_t *rows, const uint8_t *pieces, int pieces_stride,
                     const uint8_t *npieces);
/* reset (:438-449): idx i32[n] or NULL (counter-RNG draw), mask u8[n] or NULL, mode TPL_RESET_* */
int tpl_env_reset(tpl_env *e, const int32_t *idx, const uint8_t *mask, int mode, int gen_count);
/* install explicit states (tests, the single-env facade) */
int tpl_env_load(tpl_env *e, const uint16_t *rows, const uint8_t *pieces, int pieces_stride, const uint8_t *npieces,
                 const int32_t *lines, const int32_t *moves, const int8_t *st, const uint8_t *head);
/* move (:354-422): host rot/loc in, host dlines/flags/st out (outputs optional) */
int tpl_env_move(tpl_env *e, const uint8_t *rot, const uint8_t *loc, int8_t *dlines, uint8_t *flags, int8_t *st);
/* get_state (:435-436) */
int tpl_env_get_state(tpl_env *e, uint16_t *rows, uint8_t *cur, uint8_t *next, int32_t *lines, int32_t *moves,
                      int8_t *st, uint8_t *head, uint8_t *npieces, uint8_t *queue);
/* afterstates to host: feats u8[40][n][4], flags u8[40][n] */
int tpl_env_afterstates(tpl_env *e, uint8_t *feats, uint8_t *flags);
/* one host-facing rollout step: H2D actions -> tpl_step_observe (move -> auto-reset of finished envs -> afterstates of
 * the new states, one kernel) -> D2H (dlines, flags, st, feats[, aflags]).  aflags == NULL selects the compact form
 * (feats byte 0 = rows cleared | flags << 3).  Large batches are pipelined over tpl_env_chunks(e) env chunks (the D2H of a
 * chunk overlaps the kernel of the next).  feats == NULL (and aflags == NULL) leaves the features on the device, see
 * tpl_env_feats_ptr. */
int tpl_env_step_observe(tpl_env *e, const uint8_t *rot, const uint8_t *loc, int8_t *dlines, uint8_t *flags,
                         int8_t *st, uint8_t *feats, uint8_t *aflags);
/* The same step in the distinct-placements form (tpl_step_observe_distinct): only the placements that differ cross PCIe
 * (about 100 instead of 163 bytes per env-step).  rows: host u32[rows_capacity] with rows_capacity >=
 * tpl_env_distinct_capacity(e); runs: host u32[n] run descriptors (offsets into rows).  The call works through the envs in
 * tpl_env_chunks(e) chunks, each with its own region of rows, and overlaps the transfer of a chunk with the kernels of the
 * next ones; *words_copied (optional) = rows words actually transferred. */
int tpl_env_step_observe_distinct(tpl_env *e, const uint8_t *rot, const uint8_t *loc, int8_t *dlines, uint8_t *flags,
                                  int8_t *st, uint32_t *rows, int64_t rows_capacity, uint32_t *runs, int64_t *words_copied);
int64_t tpl_env_distinct_capacity(tpl_env *e);
int tpl_env_chunks(tpl_env *e);
/* pinned (page-locked) host buffers so the copies inside the calls above are true async DMA */
void *tpl_host_alloc(int64_t bytes);
void tpl_host_free(void *p);
/* device pointer of the afterstate words u8[40][n][4] written by the last tpl_env_step_observe (NULL before the first
 * one): with feats == NULL in that call the features stay in HBM for a policy that runs on the GPU and only
 * (dlines, flags, st) are copied back. */
void *tpl_env_feats_ptr(tpl_env *e);
/* raw access for callers that keep data on the device (torch): device pointer of the state planes */
void *tpl_env_state_ptr(tpl_env *e, int64_t *plane_stride);
void *tpl_env_stream(tpl_env *e);
/* kernels launched by this library in this process so far (bench.py's gpu_launches) */
long long tpl_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* TETRIS_PICLIM_H_ */
