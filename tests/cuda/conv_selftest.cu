// Standalone GPU self-test for tak_b200/csrc/conv_tc.cuh (no Python, no torch): compares the tcgen05
// implicit-GEMM conv against a naive one-thread-per-output CUDA kernel on random data, then times it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I tak_b200/csrc tests/cuda/conv_selftest.cu -o build/conv_selftest
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "conv_tc2.cuh"

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

// naive reference: out[slot][co] = bias[co] + sum_{tap,ci} in[slot+shift][ci] * w[co][tap][ci]
__global__ void conv_ref_kernel(const __nv_bfloat16* in, const __nv_bfloat16* wplain /*[co][tap][ci]*/,
                                const float* bias, float* out /*[S][128]*/, int S, int pitch, int n_boards) {
    int slot = blockIdx.x;
    int co = threadIdx.x;
    if (!conv_slot_valid(slot, pitch, n_boards)) {
        out[(size_t)slot * 128 + co] = 0.f;
        return;
    }
    float acc = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
        int s2 = slot + (tap / 3 - 1) * pitch + (tap % 3 - 1);
        for (int ci = 0; ci < 128; ++ci) {
            float a = __bfloat162float(in[((size_t)(ci >> 3) * S + s2) * 8 + (ci & 7)]);
            float b = __bfloat162float(wplain[((size_t)co * 9 + tap) * 128 + ci]);
            acc += a * b;
        }
    }
    out[(size_t)slot * 128 + co] = acc + bias[co];
}

int main(int argc, char** argv) {
    int n_boards = argc > 1 ? atoi(argv[1]) : 2000;
    int N = argc > 2 ? atoi(argv[2]) : 6;
    int pitch = N + 1, spb = pitch * pitch;
    int tiles = (n_boards * spb + pitch + 1 + CONV_TILE_M - 1) / CONV_TILE_M;
    int S = CONV_GUARD + tiles * CONV_TILE_M + CONV_GUARD;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s sms %d  boards %d N %d tiles %d S %d smem %d\n", prop.name, prop.multiProcessorCount, n_boards,
           N, tiles, S, CONV_SMEM_BYTES);

    size_t act_elems = (size_t)16 * S * 8;
    std::vector<__nv_bfloat16> h_in(act_elems, __float2bfloat16(0.f)), h_res(act_elems, __float2bfloat16(0.f));
    for (int b = 0; b < n_boards; ++b)
        for (int y = 0; y < N; ++y)
            for (int x = 0; x < N; ++x) {
                int slot = CONV_GUARD + b * spb + (y + 1) * pitch + x;
                for (int ci = 0; ci < 128; ++ci) {
                    h_in[((size_t)(ci >> 3) * S + slot) * 8 + (ci & 7)] = __float2bfloat16(urand() - 0.3f);
                    h_res[((size_t)(ci >> 3) * S + slot) * 8 + (ci & 7)] = __float2bfloat16(urand() - 0.5f);
                }
            }
    std::vector<__nv_bfloat16> h_wplain((size_t)128 * 9 * 128), h_wpacked((size_t)18 * 8 * 128 * 8),
        h_wpacked2((size_t)18 * 8 * 128 * 8);
    std::vector<float> h_bias(128);
    for (auto& v : h_bias) v = urand() - 0.5f;
    for (int co = 0; co < 128; ++co)
        for (int tap = 0; tap < 9; ++tap)
            for (int ci = 0; ci < 128; ++ci) {
                __nv_bfloat16 w = __float2bfloat16((urand() - 0.5f) * 0.06f);
                h_wplain[((size_t)co * 9 + tap) * 128 + ci] = w;
                int half = ci >> 6, kc = (ci & 63) >> 3, j = ci & 7;
                h_wpacked[((((size_t)(tap * 2 + half) * 8 + kc) * 128) + co) * 8 + j] = w;
                // CTA-pair layout: [rank = co/64][stage][kc][co%64][8]
                h_wpacked2[(((((size_t)(co >> 6) * 18 + (tap * 2 + half)) * 8 + kc) * 64) + (co & 63)) * 8 + j] = w;
            }

    __nv_bfloat16 *d_in, *d_res, *d_out, *d_wplain, *d_wpacked, *d_wpacked2;
    float *d_bias, *d_ref, *d_logits;
    CK(cudaMalloc(&d_in, act_elems * 2));
    CK(cudaMalloc(&d_res, act_elems * 2));
    CK(cudaMalloc(&d_out, act_elems * 2));
    CK(cudaMalloc(&d_wplain, h_wplain.size() * 2));
    CK(cudaMalloc(&d_wpacked, h_wpacked.size() * 2));
    CK(cudaMalloc(&d_wpacked2, h_wpacked2.size() * 2));
    CK(cudaMemcpy(d_wpacked2, h_wpacked2.data(), h_wpacked2.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_bias, 512));
    CK(cudaMalloc(&d_ref, (size_t)S * 128 * 4));
    CK(cudaMalloc(&d_logits, (size_t)128 * S * 4));
    CK(cudaMemcpy(d_in, h_in.data(), act_elems * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_res, h_res.data(), act_elems * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_wplain, h_wplain.data(), h_wplain.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_wpacked, h_wpacked.data(), h_wpacked.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_bias, h_bias.data(), 512, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0xFF, act_elems * 2));  // poison: every slot of every tile must be written
    CK(cudaMemset(d_logits, 0, (size_t)128 * S * 4));

    conv_ref_kernel<<<S, 128>>>(d_in, d_wplain, d_bias, d_ref, S, pitch, n_boards);
    CK(cudaDeviceSynchronize());
    std::vector<float> h_ref((size_t)S * 128);
    CK(cudaMemcpy(h_ref.data(), d_ref, h_ref.size() * 4, cudaMemcpyDeviceToHost));

    int fails = 0;
    const int impl_only = argc > 3 ? atoi(argv[3]) : 0;
    for (int impl = 1; impl <= 2; ++impl)
    for (int sms : {prop.multiProcessorCount, 24}) {
        if (impl_only && impl != impl_only) continue;
        for (int mode = 2; mode >= 0; --mode) {
            ConvParams p{};
            p.in = d_in; p.res = d_res; p.out = d_out; p.out_f32 = d_logits; p.w = impl == 1 ? d_wpacked : d_wpacked2;
            p.bias = d_bias;
            p.S = S; p.tiles = tiles; p.n_boards = n_boards; p.pitch = pitch; p.mode = mode;
            p.out_ch_offset = 0; p.out_ch_valid = 128;
            CK(cudaMemset(d_out, 0xFF, act_elems * 2));
            CK(cudaMemset(d_logits, 0xFF, (size_t)128 * S * 4));
            if (impl == 1) CK(conv3x3_tc_launch(p, sms, 0)); else CK(conv3x3_tc2_launch(p, sms, 0));
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("mode %d sms %d: kernel failed: %s\n", mode, sms, cudaGetErrorString(e));
                return 3;
            }
            double maxerr = 0, maxref = 0;
            long bad = 0;
            int first_bad_slot = -1, first_bad_ch = -1;
            if (mode == 2) {
                std::vector<float> h_log((size_t)128 * S);
                CK(cudaMemcpy(h_log.data(), d_logits, h_log.size() * 4, cudaMemcpyDeviceToHost));
                for (int slot = CONV_GUARD; slot < CONV_GUARD + tiles * CONV_TILE_M; ++slot)
                    for (int ch = 0; ch < 128; ++ch) {
                        float r = h_ref[(size_t)slot * 128 + ch], g = h_log[(size_t)ch * S + slot];
                        double err = fabs((double)r - g);
                        if (!(err <= 2e-3 + 1e-3 * fabs(r))) {
                            if (!bad) { first_bad_slot = slot; first_bad_ch = ch; }
                            ++bad;
                        }
                        if (err > maxerr || err != err) maxerr = err;
                        if (fabs(r) > maxref) maxref = fabs(r);
                    }
            } else {
                std::vector<__nv_bfloat16> h_out(act_elems);
                CK(cudaMemcpy(h_out.data(), d_out, act_elems * 2, cudaMemcpyDeviceToHost));
                for (int slot = CONV_GUARD; slot < CONV_GUARD + tiles * CONV_TILE_M; ++slot) {
                    bool valid = false;
                    {
                        int rel = slot - CONV_GUARD, board = rel / spb, loc = rel % spb, y = loc / pitch, x = loc % pitch;
                        valid = board < n_boards && y >= 1 && x < pitch - 1;
                    }
                    for (int ch = 0; ch < 128; ++ch) {
                        size_t off = ((size_t)(ch >> 3) * S + slot) * 8 + (ch & 7);
                        float r = h_ref[(size_t)slot * 128 + ch];
                        if (mode == 1 && valid) r += __bfloat162float(h_res[off]);
                        r = valid ? fmaxf(r, 0.f) : 0.f;
                        float g = __bfloat162float(h_out[off]);
                        double err = fabs((double)r - g);
                        if (!(err <= 2e-3 + 8e-3 * fabs(r))) {
                            if (!bad) { first_bad_slot = slot; first_bad_ch = ch; }
                            ++bad;
                        }
                        if (err > maxerr || err != err) maxerr = err;
                        if (fabs(r) > maxref) maxref = fabs(r);
                    }
                }
            }
            printf("impl %d mode %d grid<=%3d: max|err| %.3e (max|ref| %.3f) bad %ld%s\n", impl, mode, sms, maxerr, maxref, bad, bad ? "  <-- FAIL" : "  ok");
            if (bad) {
                ++fails;
                int rel = first_bad_slot - CONV_GUARD;
                printf("   first bad: slot %d (tile %d row %d board %d loc %d) ch %d\n", first_bad_slot,
                       rel / CONV_TILE_M, rel % CONV_TILE_M, rel / spb, rel % spb, first_bad_ch);
            }
        }
    }

    // ---- timing (mode 1, the res-block conv) ----
    for (int impl = 1; impl <= 2; ++impl) {
        if (impl_only && impl != impl_only) continue;
        ConvParams p{};
        p.in = d_in; p.res = d_res; p.out = d_out; p.out_f32 = d_logits; p.w = impl == 1 ? d_wpacked : d_wpacked2;
        p.bias = d_bias;
        p.S = S; p.tiles = tiles; p.n_boards = n_boards; p.pitch = pitch; p.mode = 1;
        p.out_ch_valid = 128;
        auto launch = [&]() { return impl == 1 ? conv3x3_tc_launch(p, prop.multiProcessorCount, 0)
                                               : conv3x3_tc2_launch(p, prop.multiProcessorCount, 0); };
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        for (int i = 0; i < 5; ++i) CK(launch());
        CK(cudaEventRecord(e0));
        const int reps = 50;
        for (int i = 0; i < reps; ++i) CK(launch());
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double per = ms / reps * 1e-3;
        double useful = 2.0 * n_boards * N * N * 128.0 * 1152.0;
        double issued = 2.0 * tiles * 256.0 * 128.0 * 1152.0;
        printf("timing impl %d: %.1f us/layer  useful %.1f TFLOP/s  issued %.1f TFLOP/s\n", impl, per * 1e6,
               useful / per * 1e-12, issued / per * 1e-12);
    }
    printf(fails ? "SELFTEST FAILED\n" : "SELFTEST PASSED\n");
    return fails ? 1 : 0;
}
