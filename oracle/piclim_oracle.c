/* CPU oracle (plain C) for the Tetris-piclim rollout hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * A from-scratch restatement, on 20 x 10-bit bitrows, of the reference's game/tetris.py hot path
 * (file:line citations below refer to /root/reference/game/tetris.py).  It exists to check the CUDA
 * path and to time a CPU baseline ("port") next to it; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product never links it.
 *
 * Parity status: move/get_state/prescribed-reset are PINNED -- this file is compared move by move
 * with the unmodified reference by oracle/validate_against_reference.py (>= 1e5 seeded episodes)
 * and with tests/golden/ fixtures generated from the reference.  Features (holes, bumpiness,
 * aggregate height), the flags byte and the counter-based 7-bag RNG are not in the reference:
 * PARITY UNPINNED for those definitions.
 *
 * Encoding: rows[r] bit c = column c; row 0 = top, row 19 = bottom; full row = 0x3FF.
 * state: 0 running (None), 1 won (True), 2 lost (False).
 *
 * Build: see oracle/Makefile  (gcc -O2 -fPIC -shared -pthread).
 */
#include <stdint.h>
#include <string.h>
#include <pthread.h>
#include <stdlib.h>

#define ROWS 20
#define COLS 10
#define FULL 0x3FF

enum { FLAG_TOPOUT = 1, FLAG_WIN = 2, FLAG_LOSE = 4, FLAG_ALIAS = 8, FLAG_NOPIECE = 16 };

/* ---- A. tetromino table (:23-57): row masks top->bottom, bit j = shape column j ------------ */
typedef struct { int h, w; uint8_t m[4]; int8_t prof[4]; } shape_t;
static shape_t SHAPES[7][4];
static const int NROT[7] = {2, 4, 4, 4, 2, 2, 1};
static const uint8_t MASKS[7][4][4] = {
    {{0xF,0,0,0}, {1,1,1,1}, {0}, {0}},
    {{4,7,0,0}, {3,2,2,0}, {7,1,0,0}, {1,1,3,0}},
    {{1,7,0,0}, {2,2,3,0}, {7,4,0,0}, {3,1,1,0}},
    {{2,7,0,0}, {2,3,2,0}, {7,2,0,0}, {1,3,1,0}},
    {{6,3,0,0}, {1,3,2,0}, {0}, {0}},
    {{3,6,0,0}, {2,3,1,0}, {0}, {0}},
    {{3,3,0,0}, {0}, {0}, {0}},
};
static int tables_ready = 0;

static void build_tables(void) {
    if (tables_ready) return;
    for (int p = 0; p < 7; ++p)
        for (int r = 0; r < NROT[p]; ++r) {
            shape_t *s = &SHAPES[p][r];
            s->h = 0; s->w = 0;
            for (int i = 0; i < 4; ++i) {
                s->m[i] = MASKS[p][r][i];
                if (s->m[i]) s->h = i + 1;
                for (int j = 0; j < 4; ++j) if ((s->m[i] >> j) & 1) { if (j + 1 > s->w) s->w = j + 1; }
            }
            /* bottom profile: lowest filled row (from the shape top) per shape column */
            for (int j = 0; j < 4; ++j) {
                s->prof[j] = -1;
                for (int i = 0; i < s->h; ++i) if ((s->m[i] >> j) & 1) s->prof[j] = (int8_t)i;
            }
        }
    tables_ready = 1;
}

/* Python's rot % n for any int (:61) */
static inline int pymod(int a, int n) { int r = a % n; return r < 0 ? r + n : r; }

/* ---- B. column tops and drop row (:424-433) -------------------------------------------------- */
static inline int col_top(const uint16_t *rows, int c) {
    for (int r = 0; r < ROWS; ++r) if ((rows[r] >> c) & 1) return r;
    return ROWS;
}

static int drop_row(const uint16_t *rows, const shape_t *s, int loc) {
    int best = 1000;
    for (int j = 0; j < s->w; ++j) {
        int d = col_top(rows, loc + j) - s->prof[j];
        if (d < best) best = d;
    }
    return best - 1;
}

/* ---- C. one move (:354-422).  Returns rows cleared; *topout set when the move topped out. --- */
typedef struct {
    uint16_t *rows;      /* [20] */
    const uint8_t *pieces; int npieces;
    uint8_t *head; int32_t *lines; int32_t *moves; int8_t *state;
} env_ref;

static int do_move(env_ref e, int rot, int loc, int L, int M, int *topout, int *nopiece) {
    *topout = 0; *nopiece = 0;
    if (*e.head >= e.npieces) { *nopiece = 1; return 0; }        /* reference: IndexError at :356 */
    int piece = e.pieces[*e.head]; *e.head += 1;                   /* :356 */
    const shape_t *s = &SHAPES[piece][pymod(rot, NROT[piece])];    /* :359/:61 */
    if (loc > COLS - s->w) loc = COLS - s->w;                      /* :364 */
    int d = drop_row(e.rows, s, loc);                              /* :367-369 */
    if (d < 0) { *e.state = 2; *topout = 1; return 0; }            /* :372-374 */
    for (int i = 0; i < s->h; ++i) e.rows[d + i] |= (uint16_t)(s->m[i] << loc);   /* :377-378 */
    *e.moves += 1;                                                 /* :379 */
    int k = 0; int isfull[ROWS] = {0};
    for (int i = 0; i < s->h; ++i) if (e.rows[d + i] == FULL) { isfull[d + i] = 1; ++k; }  /* :382-386 */
    if (k == 0) { if (*e.moves >= M) *e.state = 2; return 0; }     /* :389-394 */
    uint16_t tmp[ROWS]; int o = ROWS - 1;
    for (int r = ROWS - 1; r >= 0; --r) if (!isfull[r]) tmp[o--] = e.rows[r];        /* :402-405 */
    while (o >= 0) tmp[o--] = 0;                                   /* :406-407 */
    memcpy(e.rows, tmp, sizeof(tmp));
    *e.lines += k;                                                 /* :409 */
    if (*e.lines >= L) *e.state = 1;                               /* :415-417 */
    else if (*e.moves >= M) *e.state = 2;                          /* :420-422 */
    return k;
}

/* ---- F. features -------------------------------------------------------------- */
static void board_features(const uint16_t *rows, int *holes, int *bump, int *agg) {
    int h[COLS], a = 0, b = 0, cells = 0;
    for (int c = 0; c < COLS; ++c) { h[c] = ROWS - col_top(rows, c); a += h[c]; }
    for (int c = 0; c + 1 < COLS; ++c) b += abs(h[c] - h[c + 1]);
    for (int r = 0; r < ROWS; ++r) cells += __builtin_popcount(rows[r]);
    *holes = a - cells; *bump = b; *agg = a;
}

static void enumerate_afterstates(const uint16_t *rows, const uint8_t *pieces, int npieces, int head,
                                  int lines, int moves, int L, int M,
                                  uint8_t *feats /*[40][4]*/, uint8_t *flags /*[40]*/,
                                  uint16_t *boards /*[40][20] or NULL*/) {
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 10; ++c) {
            int slot = r * 10 + c;
            uint16_t b[ROWS]; memcpy(b, rows, sizeof(b));
            uint8_t hd = (uint8_t)head; int32_t ln = lines, mv = moves; int8_t st = 0;
            env_ref e = { b, pieces, npieces, &hd, &ln, &mv, &st };
            int topout, nopiece;
            int k = do_move(e, r, c, L, M, &topout, &nopiece);
            int fl = 0;
            if (nopiece) fl = FLAG_NOPIECE;
            else {
                if (topout) fl |= FLAG_TOPOUT;
                else if (k > 0 && ln >= L) fl |= FLAG_WIN;
                else if (mv >= M) fl |= FLAG_LOSE;
                int piece = pieces[head];
                const shape_t *s = &SHAPES[piece][pymod(r, NROT[piece])];
                if (r >= NROT[piece] || c > COLS - s->w) fl |= FLAG_ALIAS;
            }
            int holes = 0, bump = 0, agg = 0;
            if (!nopiece) board_features(b, &holes, &bump, &agg);
            feats[slot * 4 + 0] = (uint8_t)k; feats[slot * 4 + 1] = (uint8_t)holes;
            feats[slot * 4 + 2] = (uint8_t)bump; feats[slot * 4 + 3] = (uint8_t)agg;
            flags[slot] = (uint8_t)fl;
            if (boards) memcpy(boards + slot * ROWS, b, sizeof(b));
        }
}

/* ---- G. Philox4x32-10 + 7-bag (contract :64-108) ---------------------------------------------- */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void rng_words(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t stream, uint32_t index,
                      uint32_t out[4]) {
    uint32_t ctr[4] = { (uint32_t)env_id, (uint32_t)(env_id >> 32), episode,
                        ((stream & 0xFu) << 28) | (index & 0x0FFFFFFFu) };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    philox4x32_10(ctr, key, out);
}

static void bag_from_word(uint32_t u, uint8_t perm[7]) {
    uint32_t k = (uint32_t)(((uint64_t)u * 5040u) >> 32);
    for (int i = 0; i < 7; ++i) perm[i] = (uint8_t)i;
    for (int i = 6; i >= 1; --i) {
        uint32_t j = k % (uint32_t)(i + 1); k /= (uint32_t)(i + 1);
        uint8_t t = perm[i]; perm[i] = perm[j]; perm[j] = t;
    }
}

static void gen_pieces(uint64_t seed, uint64_t env_id, uint32_t episode, int count, uint8_t *out) {
    int n = 0, b = 0; uint32_t w[4];
    while (n < count) {
        if ((b & 3) == 0) rng_words(seed, env_id, episode, 0, (uint32_t)(b >> 2), w);
        uint8_t perm[7]; bag_from_word(w[b & 3], perm);
        for (int i = 0; i < 7 && n < count; ++i) out[n++] = perm[i];
        ++b;
    }
}

/* =================================================================================================
 * exported batch API (ctypes).  Flat arrays: rows u16[N,20], pieces u8[N,P], npieces u8[N],
 * head u8[N], lines i32[N], moves i32[N], state i8[N].
 * ================================================================================================= */
void orc_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out) { philox4x32_10(ctr, key, out); }

void orc_gen_pieces(uint64_t seed, uint64_t env_base, int n, uint32_t episode, int count, uint8_t *out /*[n,count]*/) {
    for (int i = 0; i < n; ++i) gen_pieces(seed, env_base + (uint64_t)i, episode, count, out + (size_t)i * count);
}

void orc_step_batch(int n, uint16_t *rows, const uint8_t *pieces, int P, const uint8_t *npieces,
                    uint8_t *head, int32_t *lines, int32_t *moves, int8_t *state,
                    const int32_t *rot, const int32_t *loc, int8_t *dlines, uint8_t *flags_out, int L, int M) {
    build_tables();
    for (int i = 0; i < n; ++i) {
        env_ref e = { rows + (size_t)i * ROWS, pieces + (size_t)i * P, npieces[i], head + i, lines + i, moves + i, state + i };
        int topout, nopiece;
        int k = do_move(e, rot[i], loc[i], L, M, &topout, &nopiece);
        if (dlines) dlines[i] = (int8_t)k;
        if (flags_out) flags_out[i] = (uint8_t)((topout ? FLAG_TOPOUT : 0) | (nopiece ? FLAG_NOPIECE : 0));
    }
}

void orc_features_batch(int n, const uint16_t *rows, uint8_t *out /*[n,3] holes,bump,agg*/) {
    for (int i = 0; i < n; ++i) {
        int h, b, a; board_features(rows + (size_t)i * ROWS, &h, &b, &a);
        out[i * 3] = (uint8_t)h; out[i * 3 + 1] = (uint8_t)b; out[i * 3 + 2] = (uint8_t)a;
    }
}

typedef struct {
    int lo, hi;
    const uint16_t *rows; const uint8_t *pieces; int P; const uint8_t *npieces; const uint8_t *head;
    const int32_t *lines; const int32_t *moves; int L, M; uint8_t *feats; uint8_t *flags; uint16_t *boards;
} as_job;

static void *as_worker(void *arg) {
    as_job *j = (as_job *)arg;
    for (int i = j->lo; i < j->hi; ++i)
        enumerate_afterstates(j->rows + (size_t)i * ROWS, j->pieces + (size_t)i * j->P, j->npieces[i], j->head[i],
                              j->lines[i], j->moves[i], j->L, j->M,
                              j->feats + (size_t)i * 160, j->flags + (size_t)i * 40,
                              j->boards ? j->boards + (size_t)i * 40 * ROWS : NULL);
    return NULL;
}

/* feats u8[N,40,4] (dlines, holes, bumpiness, agg height), flags u8[N,40], boards u16[N,40,20] or NULL */
void orc_afterstates_batch(int n, const uint16_t *rows, const uint8_t *pieces, int P, const uint8_t *npieces,
                           const uint8_t *head, const int32_t *lines, const int32_t *moves, int L, int M,
                           uint8_t *feats, uint8_t *flags, uint16_t *boards, int nthreads) {
    build_tables();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256]; as_job jobs[256];
    for (int t = 0; t < nthreads; ++t) {
        as_job j = { (int)((int64_t)n * t / nthreads), (int)((int64_t)n * (t + 1) / nthreads),
                     rows, pieces, P, npieces, head, lines, moves, L, M, feats, flags, boards };
        jobs[t] = j;
        if (nthreads == 1) as_worker(&jobs[t]); else pthread_create(&th[t], NULL, as_worker, &jobs[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
}

/* ---- random-agent rollout with auto-reset from a prescribed-config pool --------------------------
 * The schedule the fused CUDA rollout follows (see piclim_oracle.py: rollout_random):
 *   episode e of global env g starts from pool[mulhi(rng(seed,g,e,CONFIG,0).w0, K)];
 *   step t of that episode plays rot = w0 & 3, loc = mulhi(w1, 10) of rng(seed,g,e,ACTION,t);
 *   an env that left RUNNING (or ran out of pieces) is reset before its next move.
 * State arrays are in/out so a rollout can be continued; episode/tstep u32[N]; stats i64[8] are
 * accumulated: episodes, wins, topouts, move-limit losses, lines, moves placed, steps, resets. */
typedef struct {
    int lo, hi; uint64_t env_base, seed; int L, M;
    const uint16_t *pool_rows; const uint8_t *pool_pieces; int P; const uint8_t *pool_np; int K; int steps;
    uint16_t *rows; uint8_t *pieces; uint8_t *npieces; uint8_t *head; int32_t *lines; int32_t *moves; int8_t *state;
    uint32_t *episode; uint32_t *tstep; int64_t stats[8];
    const int32_t *weights;   /* NULL = random agent; else greedy integer linear value over the 40 afterstates */
} ro_job;

static void install(ro_job *j, int i, uint32_t ep) {
    uint32_t w[4]; rng_words(j->seed, j->env_base + (uint64_t)i, ep, 2, 0, w);
    uint32_t k = (uint32_t)(((uint64_t)w[0] * (uint32_t)j->K) >> 32);
    memcpy(j->rows + (size_t)i * ROWS, j->pool_rows + (size_t)k * ROWS, ROWS * sizeof(uint16_t));
    memcpy(j->pieces + (size_t)i * j->P, j->pool_pieces + (size_t)k * j->P, (size_t)j->P);
    j->npieces[i] = j->pool_np[k]; j->head[i] = 0; j->lines[i] = 0; j->moves[i] = 0; j->state[i] = 0;
}

static void *ro_worker(void *arg) {
    ro_job *j = (ro_job *)arg;
    for (int i = j->lo; i < j->hi; ++i) {
        for (int s = 0; s < j->steps; ++s) {
            if (j->state[i] != 0 || j->head[i] >= j->npieces[i]) {
                j->episode[i] += 1; j->tstep[i] = 0; install(j, i, j->episode[i]); j->stats[7] += 1;
            }
            int rot, loc;
            if (j->weights) {
                /* greedy: value = w0*dlines + w1*holes + w2*bump + w3*agg (+w4 win, +w5 lose/top-out);
                 * arg-max over slots, lowest slot wins ties */
                uint8_t feats[160], flags[40];
                enumerate_afterstates(j->rows + (size_t)i * ROWS, j->pieces + (size_t)i * j->P, j->npieces[i], j->head[i],
                                      j->lines[i], j->moves[i], j->L, j->M, feats, flags, NULL);
                int best = 0, bestv = 0;
                for (int s2 = 0; s2 < 40; ++s2) {
                    const int32_t *w = j->weights;
                    int v = w[0] * feats[s2 * 4] + w[1] * feats[s2 * 4 + 1] + w[2] * feats[s2 * 4 + 2] + w[3] * feats[s2 * 4 + 3];
                    if (flags[s2] & FLAG_WIN) v += w[4];
                    if (flags[s2] & (FLAG_LOSE | FLAG_TOPOUT)) v += w[5];
                    if (s2 == 0 || v > bestv) { bestv = v; best = s2; }
                }
                rot = best / 10; loc = best % 10;
            } else {
                uint32_t w[4]; rng_words(j->seed, j->env_base + (uint64_t)i, j->episode[i], 1, j->tstep[i], w);
                rot = (int)(w[0] & 3u); loc = (int)(((uint64_t)w[1] * 10u) >> 32);
            }
            env_ref e = { j->rows + (size_t)i * ROWS, j->pieces + (size_t)i * j->P, j->npieces[i], j->head + i,
                          j->lines + i, j->moves + i, j->state + i };
            int topout, nopiece; int32_t before = j->moves[i];
            int k = do_move(e, rot, loc, j->L, j->M, &topout, &nopiece);
            j->tstep[i] += 1;
            j->stats[6] += 1; j->stats[4] += k; j->stats[5] += j->moves[i] - before;
            if (j->state[i] != 0) {
                j->stats[0] += 1;
                if (j->state[i] == 1) j->stats[1] += 1; else if (topout) j->stats[2] += 1; else j->stats[3] += 1;
            }
        }
    }
    return NULL;
}

/* first_reset != 0: install episode 0 for every env before stepping (fresh rollout).
 * weights == NULL: random agent; else greedy agent with int32 weights[6]. */
void orc_rollout(int n, uint64_t env_base, uint64_t seed, int L, int M,
                        const uint16_t *pool_rows, const uint8_t *pool_pieces, int P, const uint8_t *pool_np, int K,
                        int steps, int first_reset,
                        uint16_t *rows, uint8_t *pieces, uint8_t *npieces, uint8_t *head,
                        int32_t *lines, int32_t *moves, int8_t *state, uint32_t *episode, uint32_t *tstep,
                        int64_t *stats, int nthreads, const int32_t *weights) {
    build_tables();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    ro_job *jb = (ro_job *)malloc(sizeof(ro_job) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) {
        ro_job j; memset(&j, 0, sizeof(j));
        j.lo = (int)((int64_t)n * t / nthreads); j.hi = (int)((int64_t)n * (t + 1) / nthreads);
        j.env_base = env_base; j.seed = seed; j.L = L; j.M = M;
        j.pool_rows = pool_rows; j.pool_pieces = pool_pieces; j.P = P; j.pool_np = pool_np; j.K = K; j.steps = steps;
        j.rows = rows; j.pieces = pieces; j.npieces = npieces; j.head = head; j.lines = lines; j.moves = moves;
        j.state = state; j.episode = episode; j.tstep = tstep; j.weights = weights;
        jb[t] = j;
    }
    if (first_reset)
        for (int t = 0; t < nthreads; ++t)
            for (int i = jb[t].lo; i < jb[t].hi; ++i) { episode[i] = 0; tstep[i] = 0; install(&jb[t], i, 0); }
    for (int t = 0; t < nthreads; ++t) {
        if (nthreads == 1) ro_worker(&jb[t]); else pthread_create(&th[t], NULL, ro_worker, &jb[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    for (int t = 0; t < nthreads; ++t) for (int q = 0; q < 8; ++q) stats[q] += jb[t].stats[q];
    free(jb);
}

/* Auto-reset (the TPL_RESET_DONE semantics the fused step applies after the move): every env that left RUNNING or ran out of
 * pieces starts its next episode -- counter bumped, then the same config draw as the rollouts.  Returns the number reset. */
int orc_reset_done(int n, uint64_t env_base, uint64_t seed, const uint16_t *pool_rows, const uint8_t *pool_pieces, int P,
                   const uint8_t *pool_np, int K, uint16_t *rows, uint8_t *pieces, uint8_t *npieces, uint8_t *head,
                   int32_t *lines, int32_t *moves, int8_t *state, uint32_t *episode, uint32_t *tstep) {
    ro_job j; memset(&j, 0, sizeof(j));
    j.env_base = env_base; j.seed = seed; j.pool_rows = pool_rows; j.pool_pieces = pool_pieces; j.P = P; j.pool_np = pool_np; j.K = K;
    j.rows = rows; j.pieces = pieces; j.npieces = npieces; j.head = head; j.lines = lines; j.moves = moves; j.state = state;
    int cnt = 0;
    for (int i = 0; i < n; ++i)
        if (state[i] != 0 || head[i] >= npieces[i]) {
            episode[i] += 1; if (tstep) tstep[i] = 0; install(&j, i, episode[i]); ++cnt;
        }
    return cnt;
}

int orc_abi_version(void) { return 1; }
