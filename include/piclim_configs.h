/* piclim_configs.h -- C ABI of libpiclim_carve.so: the reference's two reset-point producers, restated natively on the
 * host (no GPU involved).  They are the SUPPLY side of the reset path -- what Tetris.load_warm_reset (game/tetris.py:445-449)
 * takes from its queue -- not part of the rollout hot path and not a fallback for it.  All citations are file:line in the
 * reference tree; boards are 20 x uint16 bitrows (row 0 = top, bit c = column c), pieces are the ids of
 * game/tetris.py:8-16 (0=I 1=L 2=J 3=T 4=S 5=Z 6=O).  Every function is re-entrant; nothing is allocated for the caller.
 */
#ifndef PICLIM_CONFIGS_H
#define PICLIM_CONFIGS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- carving generator: Tetris._generate_initial_config / carve / calculate_carve (game/tetris.py:226-352),
 *      CheckpointManager (:111-137), RandomPieceGenerator (:64-108), on CPython's `random` stream ------------------- */

/* Configs for seeds seed0 .. seed0+count-1; config k == random.seed(seed0 + k); Tetris(L, M, warm_reset=False, debug=True):
 * rows[count][20], pieces[count][pieces_stride] (pieces_stride >= M + 1), npieces[count] (= M + 1, :281-284),
 * solutions[count][M][2] = the recorded (rotations, location) list in play order, -1 padded, nsol[count] (both may be null).
 * Returns 0, or -1 on bad arguments (1 <= L <= 16, M >= 1). */
int carve_generate(uint64_t seed0, int count, int L, int M, uint16_t *rows, uint8_t *pieces, int pieces_stride,
                   uint8_t *npieces, int8_t *solutions, uint8_t *nsol, int nthreads);

/* One config drawn from a caller-supplied MT19937 state (CPython's random.getstate()[1]: 624 words + index), advanced in
 * place: what Tetris(L, M, warm_reset=False) consumes from the GLOBAL stream (:226-284). */
int carve_generate_from_state(uint32_t *mt625, int L, int M, uint16_t *rows, uint8_t *pieces, int pieces_stride,
                              uint8_t *npieces, int8_t *solution, uint8_t *nsol);

/* Tetris.carve (:286-311) on a bitrow board, in place: 1 = carved, 0 = not possible, -1 = bad arguments. */
int carve_apply(uint16_t *rows, int piece, int rotations, int location, int allow_partial);

/* The first n values of random.seed(seed); [random.randint(0, hi) ...] (pins the RNG restatement in the tests). */
void carve_pyrandom_randints(uint64_t seed, int hi, int n, int32_t *out);

/* ---- forward producer: game/tetris_algo_main (TetrisGameGenerator.py, TetrisSolver.py, main.py:generate_game/solve_game),
 *      fed to the queue by forward_warm_reset_worker (game/tetris.py:482-488) ------------------------------------------ */

/* Games for seeds seed0 .. seed0+count-1: TetrisGameGenerator(seed, goal, tetrominoes, initial_height_max) and, unless
 * max_attempts < 0, TetrisSolver(board, sequence, goal, max_attempts).solve().
 * rows[count][20]; letters[count][tetrominoes] = the sequence as piece ids (piece_translations, game/tetris.py:8-16);
 * solvable[count]; failed[count] = the solver's failed_attempts; moves[count][tetrominoes][3] = its stack (name index in
 * I J L O S T Z, rotation in the generator's own table, column), -1 padded; nmoves[count].  The last four may be null. */
int forward_generate(uint64_t seed0, int count, int goal, int tetrominoes, int initial_height_max, int max_attempts,
                     uint16_t *rows, uint8_t *letters, uint8_t *solvable, int32_t *failed, int8_t *moves, uint8_t *nmoves,
                     int nthreads);

#ifdef __cplusplus
}
#endif
#endif
