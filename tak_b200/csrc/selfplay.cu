#include "engine.hpp"
namespace tb { void selfplay_destroy(tak_engine*) {} }
using namespace tb;
extern "C" {
int32_t selfplay_begin(tak_engine_t*, const tak_selfplay_config_t*) { set_error("selfplay: not built yet"); return TAK_ERR_BAD_ARG; }
int32_t selfplay_step(tak_engine_t*, int32_t, tak_selfplay_stats_t*) { set_error("selfplay: not built yet"); return TAK_ERR_BAD_ARG; }
int32_t selfplay_drain(tak_engine_t*, tak_replay_record_t*, int32_t, int32_t*) { set_error("selfplay: not built yet"); return TAK_ERR_BAD_ARG; }
}
