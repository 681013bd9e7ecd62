/* taknative.h -- C ABI of the B200-native AlphaTak self-play engine (libtaknative.so).
 *
 * This is the drop-in boundary for the ONE hot path of ViliamVadocz/tak:
 *   tak::Game::{possible_moves, play, result}   (reference: tak/src/move_gen.rs:7-102, game.rs:121-267)
 *   alpha_tak::Node::{virtual_rollout, devirtualize_path, rollout, improved_policy, pick_move, play}
 *                                               (reference: alpha-tak/src/search/mcts.rs:16-125, play.rs:13-67)
 *   alpha_tak::Network::policy_eval             (reference: alpha-tak/src/model/network.rs:34, net6.rs:124-138)
 *   train::self_play_parallel                   (reference: train/src/self_play.rs:96-262)
 * The reference has no FFI layer of its own (it is a pure-Rust workspace); each entry point below names
 * the Rust item a `extern "C"` shim crate would forward to it (INTEGRATION.md shows that shim).
 *
 * Conventions
 *   - every function returns int32_t: TAK_OK (0) or a negative tak_status code; nothing aborts.
 *   - output buffers are caller-allocated, with a capacity argument and a returned count.
 *   - a tak_engine_t owns all device memory and one CUDA stream; use one engine from one host thread at a
 *     time (engines are independent, thread-per-GPU is fine).  Calls that return host data synchronise the
 *     engine's stream before returning; calls documented "async" only enqueue work.
 *   - there is NO CPU fallback: every compute entry point fails with TAK_ERR_CUDA if the device is missing.
 *
 * Move encoding (uint16_t), shared by every entry point:
 *   bits 0-5   square index = row * N + col          (col = file a.., row = rank 1..)
 *   bits 6-7   placement: piece 0 Flat, 1 Wall, 2 Cap | spread: direction 0 Up(+) 1 Down(-) 2 Left(<) 3 Right(>)
 *   bits 8-15  spread pattern mask (0 => the move is a placement).  MSB-first: each drop of c pieces is
 *              c-1 zero bits followed by a one bit (takparse 0.5.5 `Pattern::mask`), e.g. "3a1>21" -> 0b0110_0000.
 */
#ifndef TAKNATIVE_H
#define TAKNATIVE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum tak_status {
    TAK_OK = 0,
    /* tak::PlayError (reference: tak/src/error.rs:4-15) */
    TAK_ERR_OUT_OF_BOUNDS = -1,
    TAK_ERR_ALREADY_OCCUPIED = -2,
    TAK_ERR_NO_CAPSTONE = -3,
    TAK_ERR_NO_STONES = -4,
    TAK_ERR_OPENING_NON_FLAT = -5,
    TAK_ERR_EMPTY_SQUARE = -6,
    TAK_ERR_STACK_NOT_OWNED = -7,
    TAK_ERR_STACK_WALL = -8,         /* StackError::Wall */
    TAK_ERR_STACK_CAP = -9,          /* StackError::Cap */
    TAK_ERR_TAKE_ZERO = -10,         /* TakeError::Zero */
    TAK_ERR_TAKE_CARRY_LIMIT = -11,  /* TakeError::CarryLimit */
    TAK_ERR_TAKE_STACK_SIZE = -12,   /* TakeError::StackSize */
    TAK_ERR_SPREAD_OUT_OF_BOUNDS = -13,
    /* boundary errors */
    TAK_ERR_BAD_ARG = -32,
    TAK_ERR_CUDA = -33,
    TAK_ERR_CAPACITY = -34,          /* caller buffer or device pool too small */
    TAK_ERR_PARSE = -35,
    TAK_ERR_NO_NETWORK = -36,
    TAK_ERR_INVALID_MOVE = -37       /* e.g. mcts_play of a move that is not a child (reference panics) */
} tak_status;

/* tak::GameResult (reference: tak/src/game_result.rs:4-8), one byte:
 *   0 Ongoing | 1 Winner{White} | 2 Winner{Black} | 3 Draw ; bit 4 (0x10) = road / reversible_plies flag */
enum {
    TAK_RESULT_ONGOING = 0,
    TAK_RESULT_WHITE = 1,
    TAK_RESULT_BLACK = 2,
    TAK_RESULT_DRAW = 3,
    TAK_RESULT_FLAG = 0x10
};

/* POD mirror of tak::Game<N> (reference: tak/src/game.rs:25-35, board.rs:8-10, tile.rs:7-10) for N = 3..8.
 * Square index = row * n + col.  Stack colours bottom -> top, bit i set = Black. */
typedef struct tak_state {
    uint8_t n;
    uint8_t to_move; /* 0 White, 1 Black */
    uint16_t ply;
    uint8_t white_stones, white_caps, black_stones, black_caps;
    int8_t half_komi;
    uint8_t reversible_plies;
    uint8_t _pad[6];
    uint8_t height[64];
    uint8_t top[64];        /* kind of the top piece: 0 Flat, 1 Wall, 2 Cap (0 on an empty square) */
    uint64_t stack_lo[64];  /* pieces 0..63 */
    uint64_t stack_hi[64];  /* pieces 64..127 (only reachable for N >= 7) */
} tak_state_t;

typedef struct tak_engine tak_engine_t;

typedef struct tak_engine_config {
    int32_t device;            /* CUDA ordinal */
    int32_t n;                 /* board size 3..8 */
    int32_t max_games;         /* concurrent games (and search trees) */
    int32_t nodes_per_game;    /* MCTS node-pool capacity per game and per half (0 => default) */
    int32_t max_batch;         /* max leaves per network call (0 => max_games) */
    int32_t reserved[3];
} tak_engine_config_t;

/* ---- engine ------------------------------------------------------------------------------------- */
int32_t tak_engine_create(const tak_engine_config_t* cfg, tak_engine_t** out);
int32_t tak_engine_destroy(tak_engine_t* e);
int32_t tak_engine_sync(tak_engine_t* e);
const char* tak_last_error(void);      /* thread-local text for the last failing call */
int32_t tak_version(void);

/* ---- tak::Game -----------------------------------------------------------------------------------
 * tak_games_reset        Game::with_half_komi (game.rs:57-71)              games [first, first+count)
 * tak_games_upload/...   host <-> device copy of whole states
 * tak_possible_moves     Game::possible_moves (move_gen.rs:7-30), same order; out_offsets has n+1 entries
 * tak_play               Game::play (game.rs:121-130); out_status[i] = TAK_OK or the PlayError code.  As in
 *                        the reference a failed play may leave that game in an invalid state.
 * tak_result             Game::result (game.rs:220-267)
 * tak_perft              perf_count (tak/tests/perft.rs:3-18): same counting rule, breadth-first on device  */
int32_t tak_games_reset(tak_engine_t* e, int32_t first, int32_t count, int32_t half_komi);
int32_t tak_games_upload(tak_engine_t* e, const int32_t* ids, int32_t n, const tak_state_t* states);
int32_t tak_games_download(tak_engine_t* e, const int32_t* ids, int32_t n, tak_state_t* states);
int32_t tak_possible_moves(tak_engine_t* e, const int32_t* ids, int32_t n, uint16_t* out_moves, int32_t* out_offsets,
                           int32_t cap);
int32_t tak_play(tak_engine_t* e, const int32_t* ids, const uint16_t* moves, int32_t n, int32_t* out_status);
int32_t tak_result(tak_engine_t* e, const int32_t* ids, int32_t n, uint8_t* out_results);
int32_t tak_perft(tak_engine_t* e, const tak_state_t* root, int32_t depth, uint64_t* out_nodes);
/* the same count summed over n_roots independent roots in ONE breadth-first expansion (multi-GPU perft: every rank takes a
 * share of a shallow frontier) */
int32_t tak_perft_multi(tak_engine_t* e, const tak_state_t* roots, int32_t n_roots, int32_t depth, uint64_t* out_nodes);
/* timing hook for bench.py: device milliseconds (CUDA events on the engine stream) of the last tak_perft, and
 * the number of child states it materialised / kernels it launched */
int32_t tak_perft_stats(tak_engine_t* e, double* out_ms, uint64_t* out_materialised, uint64_t* out_launches);
/* finer timing of the last tak_perft, CUDA events on the engine stream: out6 = { ms of the whole call, ms spent in the
 * expansion kernel pairs (move lists + apply/classify), child states materialised, kernels launched, children of the
 * largest single expansion, ms of that expansion } -- the roofline of movegen + play is computed from [4] and [5] */
int32_t tak_perft_profile(tak_engine_t* e, double* out6);
/* Random playouts on the device (workload A; the uniform-random playout loop of tak/tests/tps.rs:26-96
 * and symm.rs:3-68, `moves[seed % len]`, with a counter-based seed): every game in [first, first+count) is played from
 * its current position until Game::result() != Ongoing or it has added max_plies + (hash(game) % ply_spread) plies;
 * move = possible_moves()[splitmix64(seed ^ splitmix64(global_id << 32 | ply)) % len], global_id = game_id_base + slot.
 * out_plies[count] / out_results[count] may be NULL; out_totals2 = { plies played, legal moves generated } summed over
 * the games; out_ms = device milliseconds of the playout kernel. */
int32_t tak_playouts(tak_engine_t* e, int32_t first, int32_t count, uint64_t seed, int32_t game_id_base, int32_t max_plies,
                     int32_t ply_spread, int32_t* out_plies, uint8_t* out_results, uint64_t* out_totals2,
                     double* out_ms);

/* ---- host-side (cold) helpers: takparse surface ---------------------------------------------------
 * tak_move_index         alpha_tak::search::move_index (move_map.rs:19-48)
 * tak_ptn_parse/format   takparse Move FromStr / Display
 * tak_tps_format/parse   From<Game> for Tps / From<Tps> for Game (tak/src/tps.rs:7-96)                */
int32_t tak_move_index(int32_t n, uint16_t move, int32_t* out_index);
int32_t tak_policy_size(int32_t n, int32_t* out_size);
int32_t tak_ptn_parse(int32_t n, const char* text, uint16_t* out_move);
int32_t tak_ptn_format(int32_t n, uint16_t move, char* out, int32_t cap);
int32_t tak_tps_format(const tak_state_t* s, char* out, int32_t cap);
int32_t tak_tps_parse(int32_t n, const char* text, tak_state_t* out);
int32_t tak_state_init(int32_t n, int32_t half_komi, tak_state_t* out);

/* ---- alpha_tak::Network ---------------------------------------------------------------------------
 * net_create             Net5::default / Net6::default shapes (net5.rs:29-73, net6.rs:29-68); arch = 5 or 6, or
 *                        0 for the DummyNet of search/tests.rs:6-35 (policy all ones, eval 0; any N)
 * net_load_weights       fp32 host blob, tensors concatenated in the order documented in DESIGN.md
 *                        (conv weight [co][ci][3][3], conv bias, bn gamma/beta/mean/var ...); BN is folded here
 * net_load_weights_device  same blob already resident on this device (e.g. after an NCCL broadcast)
 * net_weights_size       number of fp32 elements of that blob
 * net_game_repr          alpha_tak::repr::game_repr (repr/game.rs:19-51) as fp32 [C][N][N] per state
 * net_policy_eval        Network::policy_eval (net6.rs:124-138): out_policy[B][policy_size] softmax over ALL
 *                        outputs (no legality mask), out_value[B] = tanh(fc)                              */
int32_t net_create(tak_engine_t* e, int32_t arch);
int32_t net_weights_size(tak_engine_t* e, int64_t* out_elems);
int32_t net_load_weights(tak_engine_t* e, const float* blob, int64_t elems);
int32_t net_load_weights_device(tak_engine_t* e, const void* device_blob, int64_t elems);
int32_t net_input_channels(int32_t n, int32_t* out_channels);
/* positions per 256-row tile of the conv tower (7 on 6x6, 10 on 5x5): evaluation batches that are a multiple of
 * 148 SMs x k tiles x this number fill the machine evenly (sizing hint, e.g. for the number of concurrent games) */
int32_t net_boards_per_tile(int32_t n, int32_t* out_boards);
int32_t net_game_repr(tak_engine_t* e, const tak_state_t* states, int32_t b, float* out);
int32_t net_policy_eval(tak_engine_t* e, const tak_state_t* states, int32_t b, float* out_policy, float* out_value);
/* the same forward pass, but out_logits[B][policy_size] holds the PRE-softmax policy logits (net6.rs:99-100 / net5.rs:108
 * before `softmax`): the surface the network tolerance is stated on (max abs 1e-2 against an fp32 forward) */
int32_t net_policy_logits(tak_engine_t* e, const tak_state_t* states, int32_t b, float* out_logits, float* out_value);
/* Page-locked host memory for the buffers a caller hands to the entry points above and below (cudaMallocHost /
 * cudaFreeHost): device <-> host copies into pinned memory are plain DMA at PCIe speed; into pageable memory the driver
 * stages them (measured: net_policy_eval of 32 positions, 1.16 MB of policy out: 0.69 ms into a pageable buffer). */
int32_t tak_host_alloc(size_t bytes, void** out);
int32_t tak_host_free(void* p);
/* device-resident variant used by bench.py `value`: evaluates the states of games [first, first+count) in place;
 * returns device milliseconds for `reps` forward passes */
int32_t net_forward_timed(tak_engine_t* e, int32_t first, int32_t count, int32_t reps, double* out_ms);
/* profiling variant: out[0] = ms per forward (all kernels), out[1] = ms per forward spent in the 3x3 conv kernel
 * (CUDA events around each conv launch), out[2] = conv launches per forward, out[3] = algorithmic FLOP per forward */
int32_t net_forward_profile(tak_engine_t* e, int32_t first, int32_t count, int32_t reps, double* out4);

/* ---- Network::train (alpha-tak/src/model/network.rs:37-97), Net6 and Net5 ---------------------------------------
 * net_train_begin   allocate the training state for chunks of up to max_boards positions; fp32 master weights start from
 *                   the blob last loaded with net_load_weights (VarStore of a fresh or loaded network); Adam moments 0
 * net_train_chunk   train_inner for one chunk (network.rs:59-97): inputs [b][C][n][n], pi [b][policy_size], z [b] fp32
 *                   (the tensors Example::to_tensors yields; host pointers, or device pointers with on_device = 1);
 *                   forward_training (BatchNorm on batch statistics, running statistics updated), loss_p =
 *                   -sum(pi*log_softmax)/b and loss_z = sum((z-v)^2)/b -> out_loss2, backward; gradients ACCUMULATE
 * net_train_step    opt.step() + opt.zero_grad() (network.rs:91-95): Adam(beta 0.9/0.999, eps 1e-8, L2 weight decay
 *                   folded into the gradient, as tch's nn::Adam over libtorch) on every trainable tensor
 * net_train_get     copy out the blob-shaped fp32 state: 0 weights (incl. BN running statistics), 1 gradients,
 *                   2 / 3 Adam first / second moments.  Feed `0` to net_load_weights to search with the new network.
 * net_train_grad_ptr device pointer of the gradient blob (data-parallel training: all-reduce it in place before the step)
 * (The reference's `train` binary instantiates only Net6, train/src/main.rs:42-43; `Network::train` itself is generic.)
 * The DummyNet returns TAK_ERR_BAD_ARG.                                                                              */
int32_t net_train_begin(tak_engine_t* e, int32_t max_boards);
int32_t net_train_chunk(tak_engine_t* e, const float* inputs, const float* pi, const float* z, int32_t boards,
                        int32_t on_device, float* out_loss2);
int32_t net_train_step(tak_engine_t* e, float lr, float weight_decay);
int32_t net_train_get(tak_engine_t* e, int32_t what, float* out, int64_t elems);
int32_t net_train_grad_ptr(tak_engine_t* e, void** out_device_ptr, int64_t* out_elems);
int32_t net_train_stats(tak_engine_t* e, double* out_ms_last_chunk, int32_t* out_chunks_pending, int32_t* out_steps);
int32_t net_train_end(tak_engine_t* e);

/* ---- alpha_tak::Node (one search tree per game id) --------------------------------------------------
 * mcts_tree_reset        Node::default()
 * mcts_virtual_rollout   Node::virtual_rollout x k per game (mcts.rs:26-65), leaves queued in order
 * mcts_pending           number of queued leaves + their states (what the reference hands to policy_eval)
 * mcts_devirtualize      Node::devirtualize_path for every queued leaf, in queue order (mcts.rs:67-91), with
 *                        priors/evals computed on device by the engine's network
 * mcts_devirtualize_with same, but with caller-supplied network outputs [pending][policy_size], [pending]
 * mcts_rollouts          fused loop == Node::rollout x n (mcts.rs:16-23) for every listed game in lock step
 * mcts_children          children of the root: moves (movegen order), visits, priors, expected rewards
 * mcts_root              root visits / virtual visits / expected reward
 * mcts_pick_move         Node::pick_move(exploitation=true) (play.rs:49-58): last child with max visits
 * mcts_play              Node::play (play.rs:26-43): re-root on the child (tree reuse)
 * mcts_apply_dirichlet   Node::apply_dirichlet (noise.rs:6-16) with a counter-based RNG (seeded)            */
int32_t mcts_tree_reset(tak_engine_t* e, const int32_t* ids, int32_t n);
int32_t mcts_virtual_rollout(tak_engine_t* e, const int32_t* ids, int32_t n, int32_t k);
int32_t mcts_pending(tak_engine_t* e, int32_t* out_count, int32_t* out_game_ids, tak_state_t* out_states,
                     int32_t cap);
int32_t mcts_devirtualize(tak_engine_t* e);
int32_t mcts_devirtualize_with(tak_engine_t* e, const float* policy, const float* value, int32_t count);
/* Player's pipelining (alpha-tak/src/player.rs:98-110,130-133): `rollout` queues a NEW batch of virtual rollouts before it
 * evaluates and backs up the PREVIOUS one.  mcts_reserve_pending sizes every game's leaf queue for k entries (two
 * batches); mcts_devirtualize_first backs up, for game ids[i], only its oldest counts[i] queued leaves (queue order,
 * mcts.rs:67-91) and leaves the newer ones -- and every other game's queue -- untouched. */
int32_t mcts_reserve_pending(tak_engine_t* e, int32_t k);
int32_t mcts_devirtualize_first(tak_engine_t* e, const int32_t* ids, int32_t n, const int32_t* counts);
/* Player::rollout x reps for every listed game, fused (player.rs:130-133): each repetition queues a new batch of `batch`
 * virtual rollouts per game and then evaluates + backs up the batch that was already outstanding -- with no host round
 * trip in between (the interactive analysis / playtak / pit regime). */
int32_t mcts_player_rollouts(tak_engine_t* e, const int32_t* ids, int32_t n, int32_t batch, int32_t reps);
int32_t mcts_rollouts(tak_engine_t* e, const int32_t* ids, int32_t n, int32_t n_rollouts);
int32_t mcts_children(tak_engine_t* e, int32_t id, uint16_t* out_moves, uint32_t* out_visits, float* out_priors,
                      float* out_rewards, int32_t cap, int32_t* out_count);
/* batched mcts_children: game i's children land in out_moves[i*stride ...] / out_visits[i*stride ...] */
int32_t mcts_children_batch(tak_engine_t* e, const int32_t* ids, int32_t n, uint16_t* out_moves, uint32_t* out_visits,
                            int32_t* out_counts, int32_t stride);
int32_t mcts_root(tak_engine_t* e, int32_t id, uint32_t* out_visits, uint32_t* out_virtual, float* out_reward);
/* Node::debug(depth) (alpha-tak/src/search/debug.rs:9-40): one MoveInfo per root child -- move, visits, expected
 * reward, policy and the principal continuation (pick_move(true) repeated, at most min(depth, TAK_DEBUG_MAX_DEPTH) moves,
 * each with the visit count of the node it leads to) -- sorted by descending visits as NodeDebugInfo is.  The reference
 * sorts with sort_unstable_by_key + reverse, so the order among equal visit counts is unspecified there; here equal
 * counts keep reverse move-generation order (what a stable ascending sort + reverse gives). */
#define TAK_DEBUG_MAX_DEPTH 16
typedef struct tak_move_info_t {
    uint16_t move;
    uint16_t cont_len;
    uint32_t visits;
    float reward;
    float policy;
    uint16_t cont_moves[TAK_DEBUG_MAX_DEPTH];
    uint32_t cont_visits[TAK_DEBUG_MAX_DEPTH];
} tak_move_info_t;
int32_t mcts_debug(tak_engine_t* e, int32_t id, int32_t depth, tak_move_info_t* out, int32_t cap, int32_t* out_count);

int32_t mcts_pick_move(tak_engine_t* e, const int32_t* ids, int32_t n, uint16_t* out_moves);
/* Node::pick_move(exploitation=false) (play.rs:60-65): a child drawn with probability visits / sum(visits) -- the
 * reference draws from thread_rng; here r = splitmix64(seed ^ splitmix64(game id)) % sum(visits) selects the first
 * child whose cumulative visit count exceeds r (same draw the device self-play loop makes below EXPLOIT_PLIES) */
int32_t mcts_pick_move_sampled(tak_engine_t* e, const int32_t* ids, int32_t n, uint64_t seed, uint16_t* out_moves);
int32_t mcts_play(tak_engine_t* e, const int32_t* ids, const uint16_t* moves, int32_t n);
int32_t mcts_apply_dirichlet(tak_engine_t* e, const int32_t* ids, int32_t n, float alpha, float ratio,
                             uint64_t seed);

/* ---- train::self_play_parallel ----------------------------------------------------------------------
 * One call plays `moves` lock-step plies for every game of the engine (games that end are recorded and
 * restarted, as self_play.rs:148-152,236-240 does).  Replay records are fixed-size and device-resident until
 * drained; selfplay_drain copies them out (this is the payload bench.py / the trainer all-gathers). */
typedef struct tak_selfplay_config {
    int32_t rollouts;        /* ROLLOUTS (self_play.rs:12); 800 for the headline metric */
    int32_t half_komi;       /* Game::with_komi(2) => 4 */
    int32_t instant_win;     /* 1 = "play winning moves if there are any" shortcut (self_play.rs:119-171) */
    int32_t exploit_ply;     /* EXPLOIT_PLIES (40); plies below it sample ~ visits, others argmax */
    int32_t noise_ply;       /* NOISE_PLIES (80); 0 disables Dirichlet noise (parity runs) */
    float noise_alpha;       /* 0.2 */
    float noise_ratio;       /* 0.3 */
    uint64_t seed;           /* counter-based RNG key (openings, sampling, noise) */
    int32_t max_plies;       /* safety cap per game (0 => none) */
    int32_t game_id_base;    /* global id of local game 0 (rank sharding: openings/seeds use global ids) */
    int32_t reserved[4];     /* [0] != 0: keep the positions the games hold now (mid-game starts) instead of resetting */
} tak_selfplay_config_t;

typedef struct tak_selfplay_stats {
    uint64_t plies_played;       /* searched plies (excludes the two forced opening plies) */
    uint64_t games_completed;
    uint64_t rollouts;           /* virtual rollouts started */
    uint64_t evals;              /* leaves sent to the network */
    uint64_t kernel_launches;
    uint64_t records;            /* replay records produced */
    double device_ms;            /* CUDA-event time of the whole call */
    double net_ms;               /* CUDA-event time spent in network kernels */
    uint64_t records_truncated;  /* roots with more than TAK_REPLAY_MAX_CHILDREN children (the call fails if > 0) */
    uint64_t reserved[3];
} tak_selfplay_stats_t;

/* fixed-size replay record == alpha_tak::Example (example.rs:28-33).  512 (move, visits) pairs: random play peaks at
 * 135 / 194 / 220 legal moves on 5x5 / 6x6 / 8x8.  A root with more children than that is never
 * truncated silently: selfplay_step fails with TAK_ERR_CAPACITY and tak_selfplay_stats_t::records_truncated counts it. */
#define TAK_REPLAY_MAX_CHILDREN 512
typedef struct tak_replay_record {
    int32_t game_id;            /* global game id */
    int32_t game_serial;        /* how many games this slot had completed before */
    float result;               /* +1 / 0 / -1 from the mover's perspective; NaN while the game is unfinished */
    int32_t n_children;
    tak_state_t state;
    uint16_t moves[TAK_REPLAY_MAX_CHILDREN];
    uint32_t visits[TAK_REPLAY_MAX_CHILDREN];
} tak_replay_record_t;

int32_t selfplay_begin(tak_engine_t* e, const tak_selfplay_config_t* cfg);
int32_t selfplay_step(tak_engine_t* e, int32_t moves, tak_selfplay_stats_t* out_stats);
/* selfplay_drain moves the device-side record ring to the host, completes the records of finished games (result from the
 * mover's perspective) and pops up to `cap` of them into `out`.  With out == NULL and cap == 0 it only reports how many
 * completed records are waiting (size a buffer, then call again). */
int32_t selfplay_drain(tak_engine_t* e, tak_replay_record_t* out, int32_t cap, int32_t* out_count);

/* ---- alpha_tak::Example: replay text format and 8-fold symmetry augmentation -----
 * tak_example_format     Display for Example<N> (alpha-tak/src/example.rs:81-100):
 *                        "tps;white_stones;white_caps;black_stones;black_caps;half_komi;result;mov:visits,mov:visits"
 * tak_example_parse      FromStr for Example<N> (example.rs:102-133): the game comes from the TPS, then the reserves
 *                        and half_komi fields overwrite what the TPS implied; game_id / game_serial are zeroed
 * tak_symmetry_move      Symmetry::<N>::symmetries(move)[k] (tak/src/symm.rs:41-55), k = 0..7
 * tak_symmetry_state     Symmetry::<N>::symmetries(game)[k] (symm.rs:57-97)
 * examples_to_tensors    Example::to_tensors (example.rs:63-78) for `count` examples at once, on the device: row 8e+k of
 *                        inputs [8*count][C][N][N] = game_repr(symmetries(game)[k]); of pi [8*count][policy_size] =
 *                        zeros with pi[move_index(symmetries(mov)[k])] = visits/total; of z [8*count] = result.
 *                        Host buffers, or device buffers (left on the engine's stream) when on_device != 0.            */
int32_t tak_example_format(const tak_replay_record_t* rec, char* out, int32_t cap);
int32_t tak_example_parse(int32_t n, const char* text, tak_replay_record_t* out);
int32_t tak_symmetry_move(int32_t n, uint16_t move, int32_t k, uint16_t* out);
int32_t tak_symmetry_state(const tak_state_t* s, int32_t k, tak_state_t* out);
int32_t examples_to_tensors(tak_engine_t* e, const tak_replay_record_t* recs, int32_t count, float* inputs, float* pi,
                            float* z, int32_t on_device);

/* ---- multi-GPU: NCCL inside the boundary -----------------------------------------------------
 * One engine (one process, one GPU) per rank.  Games never interact, so the rollout loop has no collective; the three
 * exchanges of the sharded loop run on the engine's stream, ordered with the kernels around them:
 * tak_comm_unique_id      ncclGetUniqueId: 128 opaque bytes made by ONE rank and handed to the others by the host program
 *                         (the Rust `train` binary would pass them over its own channel: file, socket, MPI ...)
 * tak_comm_init           ncclCommInitRank for this engine; collective over all `world` ranks
 * net_broadcast_weights   the trainer publishes a network (train/src/main.rs:101-105,120 `network = new_network`):
 *                         rank `root` passes its fp32 blob (net_load_weights layout), every rank ends with that network
 *                         loaded (BatchNorm folded, operands packed) -- ncclBroadcast over NVLink
 * selfplay_gather_replay  `examples.extend(...)` across ranks (self_play.rs:165,254): every rank passes the records it
 *                         drained and receives ALL ranks' records in rank order -- ncclAllGather of fixed-size records
 * net_train_allreduce     data-parallel Network::train: the gradient accumulators of all ranks are summed in place
 *                         (ncclAllReduce on the engine stream, i.e. after the chunks that produced them and before the
 *                         Adam kernel of net_train_step); gradients of chunks add in the reference (network.rs:84-95)
 * tak_comm_sum_u64 / tak_comm_max_f64   small host-value reductions (perft counts of a sharded frontier, timings)   */
#define TAK_COMM_ID_BYTES 128
int32_t tak_comm_unique_id(uint8_t* out, int32_t cap);
int32_t tak_comm_init(tak_engine_t* e, const uint8_t* unique_id, int32_t rank, int32_t world);
int32_t tak_comm_destroy(tak_engine_t* e);
int32_t tak_comm_info(tak_engine_t* e, int32_t* out_rank, int32_t* out_world, uint64_t* out_bytes_moved);
int32_t net_broadcast_weights(tak_engine_t* e, const float* blob, int64_t elems, int32_t root);
int32_t selfplay_gather_replay(tak_engine_t* e, const tak_replay_record_t* local, int32_t n_local,
                               tak_replay_record_t* out, int32_t cap, int32_t* out_count);
int32_t net_train_allreduce(tak_engine_t* e);
int32_t tak_comm_sum_u64(tak_engine_t* e, uint64_t* inout, int32_t n);
int32_t tak_comm_max_f64(tak_engine_t* e, double* inout, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* TAKNATIVE_H */
