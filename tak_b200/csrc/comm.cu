// NCCL inside the boundary: the three exchanges of the sharded self-play / training loop, issued
// on the ENGINE's stream so that they are ordered with the kernels around them without any host synchronisation:
//   net_broadcast_weights    rank `root`'s fp32 weight blob -> every rank's network (ncclBroadcast), the publish step of
//                            train/src/main.rs:101-105,120 (`network = new_network` / `network.save`) across GPUs
//   selfplay_gather_replay   every rank's completed replay records -> every rank (ncclAllGather of counts, then of the
//                            padded fixed-size records), the `examples.extend(...)` of self_play.rs:165,254 across GPUs
//   net_train_allreduce      sum of the ranks' fp32 gradient blobs in place (ncclAllReduce) before net_train_step:
//                            gradients of chunks ADD in the reference (network.rs:84-95), so the sum over ranks is the
//                            single-process gradient of all their chunks
// plus the two scalar reductions a sharded perft / a max-over-ranks timing needs.  The reference itself is single-process.
#include <nccl.h>

#include <algorithm>
#include <vector>

#include "engine.hpp"
#include "net.hpp"

namespace tb {

struct CommState {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    DevBuf send, recv, small;
    uint64_t bytes_moved = 0;
};

#define TB_NCCL(expr)                                                                                  \
    do {                                                                                               \
        ncclResult_t r__ = (expr);                                                                     \
        if (r__ != ncclSuccess) {                                                                      \
            tb::set_error("%s failed: %s (%s:%d)", #expr, ncclGetErrorString(r__), __FILE__, __LINE__); \
            return TAK_ERR_CUDA;                                                                       \
        }                                                                                              \
    } while (0)

void comm_destroy(tak_engine* e) {
    if (!e->comm) return;
    CommState& c = *e->comm;
    if (c.comm) ncclCommDestroy(c.comm);
    for (DevBuf* b : {&c.send, &c.recv, &c.small}) b->release();
    delete e->comm;
    e->comm = nullptr;
}

}  // namespace tb

using namespace tb;

#define TB_NEED_COMM(e)                                                                              \
    TB_CHECK((e) && (e)->comm && (e)->comm->comm, TAK_ERR_BAD_ARG, "no communicator: call tak_comm_init first"); \
    TB_CUDA(cudaSetDevice((e)->device));                                                             \
    CommState& c = *(e)->comm

extern "C" {

int32_t tak_comm_unique_id(uint8_t* out, int32_t cap) {
    static_assert(sizeof(ncclUniqueId) == TAK_COMM_ID_BYTES, "TAK_COMM_ID_BYTES must match ncclUniqueId");
    TB_CHECK(out && cap >= TAK_COMM_ID_BYTES, TAK_ERR_BAD_ARG, "tak_comm_unique_id: need %d bytes", TAK_COMM_ID_BYTES);
    ncclUniqueId id;
    TB_NCCL(ncclGetUniqueId(&id));
    std::memcpy(out, &id, sizeof(id));
    return TAK_OK;
}

int32_t tak_comm_init(tak_engine_t* e, const uint8_t* unique_id, int32_t rank, int32_t world) {
    TB_CHECK(e && unique_id && world >= 1 && rank >= 0 && rank < world, TAK_ERR_BAD_ARG, "tak_comm_init: bad argument");
    TB_CUDA(cudaSetDevice(e->device));
    comm_destroy(e);
    CommState* c = new CommState();
    e->comm = c;
    c->rank = rank;
    c->world = world;
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof(id));
    TB_NCCL(ncclCommInitRank(&c->comm, world, id, rank));
    TB_CUDA(c->small.ensure(4096));
    return TAK_OK;
}

int32_t tak_comm_destroy(tak_engine_t* e) {
    TB_CHECK(e, TAK_ERR_BAD_ARG, "null engine");
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    comm_destroy(e);
    return TAK_OK;
}

int32_t tak_comm_info(tak_engine_t* e, int32_t* out_rank, int32_t* out_world, uint64_t* out_bytes_moved) {
    TB_CHECK(e, TAK_ERR_BAD_ARG, "null engine");
    if (out_rank) *out_rank = e->comm ? e->comm->rank : 0;
    if (out_world) *out_world = e->comm ? e->comm->world : 1;
    if (out_bytes_moved) *out_bytes_moved = e->comm ? e->comm->bytes_moved : 0;
    return TAK_OK;
}

int32_t net_broadcast_weights(tak_engine_t* e, const float* blob, int64_t elems, int32_t root) {
    TB_NEED_COMM(e);
    TB_CHECK(e->net && e->net->arch != 0, TAK_ERR_NO_NETWORK, "net_broadcast_weights: no network (net_create first)");
    TB_CHECK(root >= 0 && root < c.world, TAK_ERR_BAD_ARG, "root %d out of range", root);
    TB_CHECK(elems == net_blob_elems(*e->net), TAK_ERR_BAD_ARG, "weight blob has %lld elements, architecture needs %lld",
             (long long)elems, (long long)net_blob_elems(*e->net));
    TB_CHECK(c.rank != root || blob, TAK_ERR_BAD_ARG, "the root rank must pass the blob");
    const size_t bytes = size_t(elems) * 4;
    TB_CUDA(c.send.ensure(bytes));
    if (c.rank == root) TB_CUDA(cudaMemcpyAsync(c.send.p, blob, bytes, cudaMemcpyHostToDevice, e->stream));
    TB_NCCL(ncclBroadcast(c.send.p, c.send.p, size_t(elems), ncclFloat, root, c.comm, e->stream));
    c.bytes_moved += bytes;
    std::vector<float> host(static_cast<size_t>(elems));
    TB_CUDA(cudaMemcpyAsync(host.data(), c.send.p, bytes, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    return net_load_blob(e, host.data(), elems);   // BatchNorm folding + operand packing of this rank's replica
}

int32_t selfplay_gather_replay(tak_engine_t* e, const tak_replay_record_t* local, int32_t n_local,
                               tak_replay_record_t* out, int32_t cap, int32_t* out_count) {
    TB_NEED_COMM(e);
    TB_CHECK(n_local >= 0 && (local || n_local == 0) && out_count && (out || cap == 0) && cap >= 0, TAK_ERR_BAD_ARG,
             "selfplay_gather_replay: bad argument");
    // 1. every rank's record count
    int* d_counts = c.small.as<int>();                 // [world] after the gather; my own count is staged at [512]
    TB_CHECK(c.world <= 512, TAK_ERR_BAD_ARG, "world too large");
    TB_CUDA(cudaMemcpyAsync(d_counts + 512, &n_local, 4, cudaMemcpyHostToDevice, e->stream));
    TB_NCCL(ncclAllGather(d_counts + 512, d_counts, 1, ncclInt32, c.comm, e->stream));
    std::vector<int> counts(static_cast<size_t>(c.world));
    TB_CUDA(cudaMemcpyAsync(counts.data(), d_counts, size_t(c.world) * 4, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    long long total = 0;
    int mx = 0;
    for (int v : counts) { total += v; mx = std::max(mx, v); }
    *out_count = int(total);
    if (total == 0) return TAK_OK;
    TB_CHECK(total <= cap, TAK_ERR_CAPACITY, "%lld records gathered, caller buffer holds %d", total, cap);
    // 2. the records, padded to the largest count (fixed-size records: one all-gather)
    const size_t rec = sizeof(tak_replay_record_t);
    TB_CUDA(c.send.ensure(size_t(mx) * rec));
    TB_CUDA(c.recv.ensure(size_t(mx) * rec * c.world));
    if (n_local) TB_CUDA(cudaMemcpyAsync(c.send.p, local, size_t(n_local) * rec, cudaMemcpyHostToDevice, e->stream));
    TB_NCCL(ncclAllGather(c.send.p, c.recv.p, size_t(mx) * rec, ncclChar, c.comm, e->stream));
    c.bytes_moved += size_t(mx) * rec * c.world;
    size_t at = 0;
    for (int r = 0; r < c.world; ++r) {
        if (counts[r])
            TB_CUDA(cudaMemcpyAsync(out + at, c.recv.as<uint8_t>() + size_t(r) * mx * rec, size_t(counts[r]) * rec,
                                    cudaMemcpyDeviceToHost, e->stream));
        at += size_t(counts[r]);
    }
    TB_CUDA(cudaStreamSynchronize(e->stream));
    return TAK_OK;
}

int32_t net_train_allreduce(tak_engine_t* e) {
    TB_NEED_COMM(e);
    void* grad = nullptr;
    int64_t elems = 0;
    if (int r = net_train_grad_ptr(e, &grad, &elems)) return r;
    // On the engine stream: ordered after the chunk kernels that accumulated into `grad` (net_train_chunk joins its
    // weight-gradient stream before it returns) and before the Adam kernel of net_train_step.
    TB_NCCL(ncclAllReduce(grad, grad, size_t(elems), ncclFloat, ncclSum, c.comm, e->stream));
    c.bytes_moved += size_t(elems) * 4;
    return TAK_OK;
}

int32_t tak_comm_sum_u64(tak_engine_t* e, uint64_t* inout, int32_t n) {
    TB_NEED_COMM(e);
    TB_CHECK(inout && n >= 0 && n <= 256, TAK_ERR_BAD_ARG, "tak_comm_sum_u64: bad argument");
    if (n == 0) return TAK_OK;
    TB_CUDA(cudaMemcpyAsync(c.small.p, inout, size_t(n) * 8, cudaMemcpyHostToDevice, e->stream));
    TB_NCCL(ncclAllReduce(c.small.p, c.small.p, size_t(n), ncclUint64, ncclSum, c.comm, e->stream));
    TB_CUDA(cudaMemcpyAsync(inout, c.small.p, size_t(n) * 8, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    return TAK_OK;
}

int32_t tak_comm_max_f64(tak_engine_t* e, double* inout, int32_t n) {
    TB_NEED_COMM(e);
    TB_CHECK(inout && n >= 0 && n <= 256, TAK_ERR_BAD_ARG, "tak_comm_max_f64: bad argument");
    if (n == 0) return TAK_OK;
    TB_CUDA(cudaMemcpyAsync(c.small.p, inout, size_t(n) * 8, cudaMemcpyHostToDevice, e->stream));
    TB_NCCL(ncclAllReduce(c.small.p, c.small.p, size_t(n), ncclDouble, ncclMax, c.comm, e->stream));
    TB_CUDA(cudaMemcpyAsync(inout, c.small.p, size_t(n) * 8, cudaMemcpyDeviceToHost, e->stream));
    TB_CUDA(cudaStreamSynchronize(e->stream));
    return TAK_OK;
}

}  // extern "C"
