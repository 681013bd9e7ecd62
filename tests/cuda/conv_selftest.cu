// Standalone GPU self-test for tak_b200/csrc/conv_tc3.cuh (no Python, no torch): compares the tcgen05 implicit-GEMM
// conv against a naive dense NHWC convolution (one thread per output, layout-independent: it knows nothing about
// strips, pad columns or halos) on random data, checks that every pad slot is written as zero, checks the per-slot
// softmax partials of the policy mode, then times the kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I tak_b200/csrc tests/cuda/conv_selftest.cu -o build/conv_selftest
// Usage: conv_selftest [boards=2000] [N=6] [slabs=8] [pad_free=0]   (pad_free = 1: the inference tower's strip without pad
//        columns, conv3x3_tc3_kernel<false, true>: 7 boards of 6x6 per tile, masked slab copies for the horizontal taps)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "conv_tc3.cuh"

using namespace tb;

#define CK(x)                                                                           \
    do {                                                                                \
        cudaError_t e_ = (x);                                                           \
        if (e_ != cudaSuccess) {                                                        \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                    \
        }                                                                               \
    } while (0)

static uint64_t rng_state = 0x1234567ULL;
static inline uint64_t splitmix() {
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline float urand() { return (splitmix() >> 40) * (1.0f / 16777216.0f); }

// naive reference on dense tensors: in [b][y][x][ci] bf16, w [co][tap][ci] bf16 -> out [b][y][x][co] fp32
__global__ void conv_ref_kernel(const __nv_bfloat16* in, const __nv_bfloat16* w, const float* bias, float* out, int N,
                                int c_in) {
    const int pos = blockIdx.x;  // b*N*N + y*N + x
    const int co = threadIdx.x;
    const int b = pos / (N * N), y = (pos / N) % N, x = pos % N;
    float acc = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy < 0 || yy >= N || xx < 0 || xx >= N) continue;
        const __nv_bfloat16* a = in + ((size_t)(b * N + yy) * N + xx) * 128;
        for (int ci = 0; ci < c_in; ++ci)
            acc += __bfloat162float(a[ci]) * __bfloat162float(w[((size_t)co * 9 + tap) * 128 + ci]);
    }
    out[(size_t)pos * 128 + co] = acc + bias[co];
}

int main(int argc, char** argv) {
    const int n_boards = argc > 1 ? atoi(argv[1]) : 2000;
    const int N = argc > 2 ? atoi(argv[2]) : 6;
    const int slabs = argc > 3 ? atoi(argv[3]) : 8;
    const bool pad_free = argc > 4 && atoi(argv[4]) != 0;
    if (N != 5 && N != 6) { printf("N must be 5 or 6\n"); return 2; }
    ConvParams lay{};
    conv_params_set_layout(lay, N, pad_free);
    auto launch = [&](const ConvParams& q, int sms, cudaStream_t st) {
        return pad_free ? conv3x3_tc3_launch<false, true>(q, sms, st) : conv3x3_tc3_launch<false, false>(q, sms, st);
    };
    const int tiles = ((n_boards + lay.bpt - 1) / lay.bpt + C3_TILE_ALIGN - 1) / C3_TILE_ALIGN * C3_TILE_ALIGN;
    const int S = tiles * C3_TILE_M;
    const int c_in = slabs * 16;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s sms %d  boards %d N %d bpt %d pitch %d tiles %d S %d slabs %d smem %d pad_free %d\n", prop.name,
           prop.multiProcessorCount, n_boards, N, lay.bpt, lay.pitch, tiles, S, slabs,
           pad_free ? C3_PF_SMEM_BYTES : C3_SMEM_BYTES, int(pad_free));
    auto slot_of = [&](int b, int y, int x) { return (size_t)(b / lay.bpt) * 256 + y * lay.pitch + (b % lay.bpt) * lay.bw + x; };

    const size_t act_elems = (size_t)16 * S * 8;
    const size_t dense = (size_t)n_boards * N * N * 128;
    std::vector<__nv_bfloat16> h_in(act_elems, __float2bfloat16(0.f)), h_res(act_elems, __float2bfloat16(0.f));
    std::vector<__nv_bfloat16> h_din(dense, __float2bfloat16(0.f)), h_dres(dense, __float2bfloat16(0.f));
    for (int b = 0; b < n_boards; ++b)
        for (int y = 0; y < N; ++y)
            for (int x = 0; x < N; ++x) {
                const size_t slot = slot_of(b, y, x), pos = ((size_t)b * N + y) * N + x;
                for (int ci = 0; ci < 128; ++ci) {
                    // channels beyond c_in hold junk on purpose: with slabs < 8 the kernel must never read them
                    const __nv_bfloat16 a = __float2bfloat16(urand() - 0.3f), r = __float2bfloat16(urand() - 0.5f);
                    h_in[((size_t)(ci >> 3) * S + slot) * 8 + (ci & 7)] = a;
                    h_res[((size_t)(ci >> 3) * S + slot) * 8 + (ci & 7)] = r;
                    h_din[pos * 128 + ci] = a;
                    h_dres[pos * 128 + ci] = r;
                }
            }
    std::vector<__nv_bfloat16> h_wplain((size_t)128 * 9 * 128), h_wpacked(C3_W_LAYER_ELEMS, __float2bfloat16(0.f));
    std::vector<float> h_bias(128);
    for (auto& v : h_bias) v = urand() - 0.5f;
    for (int co = 0; co < 128; ++co)
        for (int tap = 0; tap < 9; ++tap)
            for (int ci = 0; ci < 128; ++ci) {
                const __nv_bfloat16 w = __float2bfloat16((urand() - 0.5f) * 0.06f);
                h_wplain[((size_t)co * 9 + tap) * 128 + ci] = w;
                const int slab = ci >> 4, kc = (ci >> 3) & 1, j = ci & 7;
                h_wpacked[(((((size_t)slab * 9 + tap) * 2 + kc) * 128) + co) * 8 + j] = w;
            }

    __nv_bfloat16 *d_in, *d_res, *d_out, *d_wplain, *d_wpacked, *d_din;
    float *d_bias, *d_ref, *d_logits;
    float2* d_part;
    CK(cudaMalloc(&d_in, act_elems * 2));
    CK(cudaMalloc(&d_res, act_elems * 2));
    CK(cudaMalloc(&d_out, act_elems * 2));
    CK(cudaMalloc(&d_din, dense * 2));
    CK(cudaMalloc(&d_wplain, h_wplain.size() * 2));
    CK(cudaMalloc(&d_wpacked, h_wpacked.size() * 2));
    CK(cudaMalloc(&d_bias, 512));
    CK(cudaMalloc(&d_ref, dense * 4));
    CK(cudaMalloc(&d_logits, (size_t)128 * S * 4));
    CK(cudaMalloc(&d_part, (size_t)4 * S * 8));
    CK(cudaMemcpy(d_in, h_in.data(), act_elems * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_res, h_res.data(), act_elems * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_din, h_din.data(), dense * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_wplain, h_wplain.data(), h_wplain.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_wpacked, h_wpacked.data(), h_wpacked.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_bias, h_bias.data(), 512, cudaMemcpyHostToDevice));

    conv_ref_kernel<<<n_boards * N * N, 128>>>(d_din, d_wplain, d_bias, d_ref, N, c_in);
    CK(cudaDeviceSynchronize());
    std::vector<float> h_ref(dense);
    CK(cudaMemcpy(h_ref.data(), d_ref, dense * 4, cudaMemcpyDeviceToHost));

    auto layer = [&](const __nv_bfloat16* in, const __nv_bfloat16* res, __nv_bfloat16* out, int mode, int ch_valid) {
        ConvLayerDesc d{};
        d.in = in; d.res = res; d.out = out; d.out_f32 = d_logits; d.partials = d_part; d.w = d_wpacked;
        d.bias = d_bias; d.slabs = slabs; d.mode = mode; d.out_ch_offset = 0; d.out_ch_valid = ch_valid; d.group = 0;
        return d;
    };
    auto params = [&](int mode, int ch_valid) {
        ConvParams p = lay;
        p.S = S; p.tile_begin = 0; p.tile_end = tiles; p.n_boards = n_boards;
        p.n_layers = 1;
        p.layers[0] = layer(d_in, d_res, d_out, mode, ch_valid);
        return p;
    };

    int fails = 0;
    for (int sms : {prop.multiProcessorCount, 24}) {
        for (int mode = 2; mode >= 0; --mode) {
            const int ch_valid = mode == 2 ? 123 : 128;
            ConvParams p = params(mode, ch_valid);
            CK(cudaMemset(d_out, 0xFF, act_elems * 2));  // poison: every slot of every tile must be written
            CK(cudaMemset(d_logits, 0xFF, (size_t)128 * S * 4));
            CK(cudaMemset(d_part, 0xFF, (size_t)4 * S * 8));
            CK(launch(p, sms, 0));
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("mode %d sms %d: kernel failed: %s\n", mode, sms, cudaGetErrorString(e));
                return 3;
            }
            double maxerr = 0, maxref = 0;
            long bad = 0, bad_pad = 0, bad_part = 0;
            // which slots are real squares
            std::vector<int> pos_of(S, -1);
            for (int b = 0; b < n_boards; ++b)
                for (int y = 0; y < N; ++y)
                    for (int x = 0; x < N; ++x) pos_of[slot_of(b, y, x)] = (b * N + y) * N + x;
            if (mode == 2) {
                std::vector<float> h_log((size_t)128 * S);
                std::vector<float2> h_part((size_t)4 * S);
                CK(cudaMemcpy(h_log.data(), d_logits, h_log.size() * 4, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(h_part.data(), d_part, (size_t)4 * S * 8, cudaMemcpyDeviceToHost));
                for (int slot = 0; slot < S; ++slot) {
                    double mx = -1e30, sum = 0;
                    for (int ch = 0; ch < ch_valid; ++ch) {
                        const float g = h_log[(size_t)ch * S + slot];
                        if (pos_of[slot] < 0) continue;  // logits of pad slots are never read
                        const float r = h_ref[(size_t)pos_of[slot] * 128 + ch];
                        const double err = fabs((double)r - g);
                        if (!(err <= 2e-3 + 1e-3 * fabs(r))) ++bad;
                        if (err > maxerr || err != err) maxerr = err;
                        if (fabs(r) > maxref) maxref = fabs(r);
                        if (g > mx) mx = g;
                    }
                    if (pos_of[slot] < 0) continue;
                    for (int ch = 0; ch < ch_valid; ++ch) sum += exp((double)h_log[(size_t)ch * S + slot] - mx);
                    // the epilogue leaves one partial per 32-channel lane quarter: merge the four
                    double pm = -1e30, ps = 0;
                    for (int qq = 0; qq < 4; ++qq) pm = fmax(pm, (double)h_part[(size_t)qq * S + slot].x);
                    for (int qq = 0; qq < 4; ++qq) {
                        const float2 pr = h_part[(size_t)qq * S + slot];
                        if (pr.y > 0) ps += pr.y * exp((double)pr.x - pm);
                    }
                    if (!((float)pm == (float)mx) || !(fabs(ps - sum) <= 1e-4 * sum)) ++bad_part;
                }
            } else {
                std::vector<__nv_bfloat16> h_out(act_elems);
                CK(cudaMemcpy(h_out.data(), d_out, act_elems * 2, cudaMemcpyDeviceToHost));
                for (int slot = 0; slot < S; ++slot)
                    for (int ch = 0; ch < 128; ++ch) {
                        const size_t off = ((size_t)(ch >> 3) * S + slot) * 8 + (ch & 7);
                        const float g = __bfloat162float(h_out[off]);
                        if (pos_of[slot] < 0) { if (g != 0.f || g != g) ++bad_pad; continue; }
                        float r = h_ref[(size_t)pos_of[slot] * 128 + ch];
                        if (mode == 1) r += __bfloat162float(h_dres[(size_t)pos_of[slot] * 128 + ch]);
                        r = fmaxf(r, 0.f);
                        const double err = fabs((double)r - g);
                        if (!(err <= 2e-3 + 8e-3 * fabs(r))) ++bad;
                        if (err > maxerr || err != err) maxerr = err;
                        if (fabs(r) > maxref) maxref = fabs(r);
                    }
            }
            const bool ok = !bad && !bad_pad && !bad_part;
            printf("mode %d grid<=%3d: max|err| %.3e (max|ref| %.3f) bad %ld bad_pad %ld bad_partials %ld%s\n", mode,
                   sms, maxerr, maxref, bad, bad_pad, bad_part, ok ? "  ok" : "  <-- FAIL");
            if (!ok) ++fails;
        }
    }

    // ---- fused tower: three chained layers in ONE launch must equal three single-layer launches bit for bit ----
    if (slabs == 8) {
        __nv_bfloat16 *d_a, *d_b, *d_c1, *d_c2;
        CK(cudaMalloc(&d_a, act_elems * 2));
        CK(cudaMalloc(&d_b, act_elems * 2));
        CK(cudaMalloc(&d_c1, act_elems * 2));
        CK(cudaMalloc(&d_c2, act_elems * 2));
        for (int sms : {prop.multiProcessorCount, 24, 5}) {
            ConvParams q = lay;
            q.S = S; q.tile_begin = 0; q.tile_end = tiles; q.n_boards = n_boards;
            const ConvLayerDesc l0 = layer(d_in, nullptr, d_a, 0, 128), l1 = layer(d_a, nullptr, d_b, 0, 128);
            for (int pass = 0; pass < 2; ++pass) {
                __nv_bfloat16* d_c = pass ? d_c2 : d_c1;
                CK(cudaMemset(d_a, 0xFF, act_elems * 2));
                CK(cudaMemset(d_b, 0xFF, act_elems * 2));
                CK(cudaMemset(d_c, 0xFF, act_elems * 2));
                const ConvLayerDesc l2 = layer(d_b, d_a, d_c, 1, 128);
                if (pass == 0) {
                    for (const ConvLayerDesc& l : {l0, l1, l2}) {
                        q.n_layers = 1; q.layers[0] = l;
                        CK(launch(q, sms, 0));
                    }
                } else {
                    q.n_layers = 3; q.layers[0] = l0; q.layers[1] = l1; q.layers[2] = l2;
                    CK(launch(q, sms, 0));
                }
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("tower pass %d sms %d: kernel failed: %s\n", pass, sms, cudaGetErrorString(e)); return 3; }
            }
            std::vector<__nv_bfloat16> h1(act_elems), h2(act_elems);
            CK(cudaMemcpy(h1.data(), d_c1, act_elems * 2, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(h2.data(), d_c2, act_elems * 2, cudaMemcpyDeviceToHost));
            const bool same = memcmp(h1.data(), h2.data(), act_elems * 2) == 0;
            double sum = 0;
            for (size_t i = 0; i < act_elems; i += 97) sum += __bfloat162float(h2[i]);
            printf("tower 3 layers fused vs separate, grid<=%3d: %s (checksum %.3f)\n", sms, same ? "identical  ok" : "DIFFERENT  <-- FAIL", sum);
            if (!same) ++fails;
        }
    }

    // ---- timing (mode 1, the res-block conv) ----
    {
        ConvParams p = params(1, 128);
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        for (int i = 0; i < 5; ++i) CK(launch(p, prop.multiProcessorCount, 0));
        CK(cudaEventRecord(e0));
        const int reps = 50;
        for (int i = 0; i < reps; ++i) CK(launch(p, prop.multiProcessorCount, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double per = ms / reps * 1e-3;
        const double useful = 2.0 * n_boards * N * N * 128.0 * 9.0 * c_in;
        const double issued = 2.0 * tiles * 256.0 * 128.0 * 9.0 * c_in;
        printf("timing: %.1f us/layer  useful %.1f TFLOP/s  issued %.1f TFLOP/s\n", per * 1e6, useful / per * 1e-12,
               issued / per * 1e-12);
        // a 32-layer tower in one launch (ping-pong between two buffers; values are irrelevant here)
        ConvParams q = p;
        q.n_layers = 32;
        for (int l = 0; l < 32; ++l) q.layers[l] = layer(l & 1 ? d_out : d_in, nullptr, l & 1 ? d_in : d_out, 0, 128);
        for (int i = 0; i < 2; ++i) CK(launch(q, prop.multiProcessorCount, 0));
        CK(cudaEventRecord(e0));
        const int reps2 = 5;
        for (int i = 0; i < reps2; ++i) CK(launch(q, prop.multiProcessorCount, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double per2 = ms / reps2 / 32 * 1e-3;
        printf("timing, 32-layer tower in one launch: %.1f us/layer  useful %.1f TFLOP/s  issued %.1f TFLOP/s\n",
               per2 * 1e6, useful / per2 * 1e-12, issued / per2 * 1e-12);
    }
    printf(fails ? "SELFTEST FAILED\n" : "SELFTEST PASSED\n");
    return fails ? 1 : 0;
}
