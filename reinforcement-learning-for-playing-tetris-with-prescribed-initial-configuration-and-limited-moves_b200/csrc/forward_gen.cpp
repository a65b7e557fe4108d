// forward_gen.cpp -- native restatement of the reference's FORWARD reset-point producer (host code, g++):
// game/tetris_algo_main/TetrisGameGenerator.py (seeded random stacking + 7-bag sequence), TetrisSolver.py (the greedy
// depth-first solvability filter) and main.py:generate_batch (seeds 0..99, keep the solvable games), i.e. what
// forward_warm_reset_worker (game/tetris.py:482-488) feeds into the reset queue.  Like carve_gen.cpp this is the supply
// side of the reset path, not the rollout path, and not a fallback for it.
//
// Two things differ from game/tetris.py and are restated as they are:
//  * the drop: a shape starts INSIDE the board at row 0 and slides down while it does not overlap
//    (TetrisGameGenerator.place_tetromino :44-51), not the column-top rule of Tetris.move; every full row of the board
//    is cleared (:53-56), not only the piece's rows;
//  * the orientation tables (:6-14) list the rotations in another order than game/tetris.py:23-57; only boards and
//    piece LETTERS leave this module (translate, game/tetris.py:8-20), so the order never meets the rollout path.
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "pyrandom.h"

namespace {

using tplgen::PyRandom;

// tetromino_shapes (TetrisGameGenerator.py:6-14 == TetrisSolver.py:5-13); names in the generator's order I J L O S T Z
// (:24); rows top -> bottom, bit j = column j
struct FShape { int rows, cols; uint8_t m[4]; };
const int F_NROT[7] = {2, 4, 4, 1, 2, 4, 2};
const FShape F_SHAPES[7][4] = {
    /* I */ {{1, 4, {0xF}}, {4, 1, {1, 1, 1, 1}}},
    /* J */ {{2, 3, {1, 7}}, {3, 2, {3, 1, 1}}, {2, 3, {7, 4}}, {3, 2, {2, 2, 3}}},
    /* L */ {{2, 3, {4, 7}}, {3, 2, {1, 1, 3}}, {2, 3, {7, 1}}, {3, 2, {3, 2, 2}}},
    /* O */ {{2, 2, {3, 3}}},
    /* S */ {{2, 3, {6, 3}}, {3, 2, {1, 3, 2}}},
    /* T */ {{2, 3, {2, 7}}, {3, 2, {1, 3, 1}}, {2, 3, {7, 2}}, {3, 2, {2, 3, 2}}},
    /* Z */ {{2, 3, {3, 6}}, {3, 2, {2, 3, 1}}},
};
// piece_translations (game/tetris.py:8-16): generator name index (I J L O S T Z) -> id of game/tetris.py (I L J T S Z O)
const uint8_t F_TO_ID[7] = {0, 2, 1, 6, 4, 3, 5};

constexpr int H = 20, W = 10;

struct FBoard {
    uint16_t r[H];
    bool overlaps(const FShape &s, int row, int col) const {
        for (int i = 0; i < s.rows; ++i) if (r[row + i] & ((uint16_t)s.m[i] << col)) return true;
        return false;
    }
    // is_valid_move (:31-42)
    bool valid(const FShape &s, int row, int col) const {
        if (row + s.rows > H || col < 0 || col + s.cols > W) return false;
        return !overlaps(s, row, col);
    }
    // calculate_placement_height (:59-67): rows the shape can slide from row 0 before it overlaps or leaves the board
    int placement_height(const FShape &s, int col) const {
        int h = 0;
        while (h + s.rows <= H && !overlaps(s, h, col)) ++h;
        return h;
    }
    // place_tetromino (:44-51) + clear_lines (:53-56); returns the number of cleared rows
    int place(const FShape &s, int col) {
        int row = 0;
        while (row + s.rows <= H && !overlaps(s, row, col)) ++row;
        for (int i = 0; i < s.rows; ++i) r[row - 1 + i] |= (uint16_t)s.m[i] << col;
        int k = 0, w = H - 1;
        for (int i = H - 1; i >= 0; --i) { if (r[i] == 0x3FF) ++k; else r[w--] = r[i]; }
        while (w >= 0) r[w--] = 0;
        return k;
    }
};

// TetrisGameGenerator.__init__ (:15-28): random.seed(seed); fill_grid(); generate_tetromino_sequence(tetrominoes)
void forward_game(uint64_t seed, int tetrominoes, int initial_height_max, FBoard &b, std::vector<uint8_t> &seq /* name indices */) {
    PyRandom rng; rng.seed(seed);
    std::memset(b.r, 0, sizeof(b.r));
    for (;;) {                                                                  // fill_grid :70-83
        const int t = (int)rng.randbelow(7);                                    // random.choice(names)
        const int rot = rng.randint(0, F_NROT[t] - 1);
        const FShape &s = F_SHAPES[t][rot];
        const int col = rng.randint(0, W - s.cols);
        if (b.valid(s, 0, col)) {
            if (H + 1 - b.placement_height(s, col) <= initial_height_max) b.place(s, col);
            else break;
        }
    }
    seq.clear();                                                                // generate_tetromino_sequence :88-104
    while ((int)seq.size() < tetrominoes || seq.empty()) {
        uint8_t bag[7] = {0, 1, 2, 3, 4, 5, 6};
        rng.shuffle(bag, 7);                // (the reshuffle condition at :97 compares neighbours of a permutation: never true)
        seq.insert(seq.end(), bag, bag + 7);
        if ((int)seq.size() >= tetrominoes) break;
    }
    if (tetrominoes > 0) seq.resize((size_t)tetrominoes);                       // sequence[:max_moves] if max_moves else sequence
}

// TetrisSolver (TetrisSolver.py:17-163)
struct Solver {
    FBoard board;
    const uint8_t *seq; int nseq, pos;          // the deque: seq[pos..nseq) is what is left
    int lines, failed, goal, max_attempts;
    std::vector<int8_t> stack;                  // (piece name index, rotation, column) triples

    // evaluate_columns(...)[:1] (:95-104, :117): the column with the greatest placement height, lowest index on ties
    int best_column(const FShape &s) const {
        int best = 0, bh = -1;
        for (int c = 0; c <= W - s.cols; ++c) { const int h = board.placement_height(s, c); if (h > bh) { bh = h; best = c; } }
        return best;
    }
    bool game_over() const { return board.r[0] != 0; }                           // :92-93

    bool solve(int current) {                                                    // :112-163 (current already popped)
        for (int rot = 0; rot < F_NROT[current]; ++rot) {
            const FShape &s = F_SHAPES[current][rot];
            const int col = best_column(s);
            if (failed >= max_attempts) return false;                            // :120-122
            const FBoard copy = board; const int lines0 = lines;
            if (board.valid(s, 0, col)) lines += board.place(s, col);            // :126-127
            else { ++failed; continue; }                                         // :128-130
            if (game_over()) { board = copy; lines = lines0; ++failed; continue; }   // :132-136
            else if (lines >= goal) { push(current, rot, col); return true; }    // :138-140
            else if (pos < nseq) {                                               // :142-151
                push(current, rot, col);
                const int next = seq[pos++];
                if (solve(next)) return true;
                --pos; stack.resize(stack.size() - 3);
                lines = lines0; board = copy;
            } else { board = copy; lines = lines0; ++failed; }                   // :153-156
            // :158-161  `len(current)` is the length of the one-letter NAME, i.e. the test is rotation == 0
            if (rot == 0 && col == W - s.cols) { ++failed; board = copy; lines = lines0; }
        }
        return false;
    }
    void push(int t, int rot, int col) { stack.push_back((int8_t)t); stack.push_back((int8_t)rot); stack.push_back((int8_t)col); }
};

bool solve_game(const FBoard &b, const std::vector<uint8_t> &seq, int goal, int max_attempts, int *failed, std::vector<int8_t> *stack) {
    Solver sv; sv.board = b; sv.seq = seq.data(); sv.nseq = (int)seq.size(); sv.pos = 0;
    sv.lines = 0; sv.failed = 0; sv.goal = goal; sv.max_attempts = max_attempts;
    bool ok = false;
    if (sv.nseq > 0) { const int first = sv.seq[sv.pos++]; ok = sv.solve(first); }
    if (failed) *failed = sv.failed;
    if (stack) *stack = sv.stack;
    return ok;
}

}  // namespace

extern "C" {

// Games for seeds seed0 .. seed0+count-1: TetrisGameGenerator(seed, goal, tetrominoes, initial_height_max) followed by
// TetrisSolver(board, sequence, goal, max_attempts).solve()  (main.py:generate_game / solve_game).
//   rows u16[count][20] (row 0 = top, bit c = column c);  letters u8[count][tetrominoes]: the sequence as ids of
//   game/tetris.py (piece_translations);  solvable u8[count];  failed i32[count] (failed_attempts, may be null);
//   moves i8[count][tetrominoes][3] = the solver's stack (name index I J L O S T Z, rotation, column), -1 padded, and
//   nmoves u8[count] (both may be null).  max_attempts < 0: generate only.
int forward_generate(uint64_t seed0, int count, int goal, int tetrominoes, int initial_height_max, int max_attempts,
                     uint16_t *rows, uint8_t *letters, uint8_t *solvable, int32_t *failed, int8_t *moves, uint8_t *nmoves,
                     int nthreads) {
    if (count < 0 || tetrominoes < 1 || tetrominoes > 255 || !rows || !letters) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > count) nthreads = count > 0 ? count : 1;
    auto work = [&](int t) {
        for (int k = t; k < count; k += nthreads) {
            FBoard b; std::vector<uint8_t> seq;
            forward_game(seed0 + (uint64_t)k, tetrominoes, initial_height_max, b, seq);
            std::memcpy(rows + (size_t)k * 20, b.r, sizeof(b.r));
            for (int p = 0; p < tetrominoes; ++p) letters[(size_t)k * tetrominoes + p] = F_TO_ID[seq[p]];
            if (max_attempts >= 0) {
                int f = 0; std::vector<int8_t> st;
                const bool ok = solve_game(b, seq, goal, max_attempts, &f, &st);
                if (solvable) solvable[k] = ok ? 1 : 0;
                if (failed) failed[k] = f;
                if (moves) {
                    int8_t *mv = moves + (size_t)k * tetrominoes * 3;
                    std::memset(mv, -1, (size_t)tetrominoes * 3);
                    if (ok) std::memcpy(mv, st.data(), st.size());
                }
                if (nmoves) nmoves[k] = ok ? (uint8_t)(st.size() / 3) : 0;
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    return 0;
}

}  // extern "C"
