// Standalone GPU self-test for tak_b200/csrc/wgrad_tc.cuh: the tcgen05 weight-gradient GEMM (MN-major operands, tap
// shifts over zero halos, split-K over tiles) against a naive one-thread-per-weight reference on random strip planes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I tak_b200/csrc tests/cuda/wgrad_selftest.cu -o build/wgrad_selftest
// Usage: wgrad_selftest [tiles=300] [c_in=128] [pitch=42]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "wgrad_tc.cuh"

using namespace tb;

#define CK(x)                                                                               \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                        \
        }                                                                                   \
    } while (0)

static uint64_t rng_state = 0x2468ACEULL;
static inline uint64_t splitmix() {
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline float urand() { return (splitmix() >> 40) * (1.0f / 16777216.0f); }

// planes [chunk][S][8]; ref[(co*c_in + ci)*9 + tap] = sum_tile sum_s dy[s][co] * x[s + shift][ci] (zero outside the tile)
__global__ void wgrad_ref_kernel(const __nv_bfloat16* dy, const __nv_bfloat16* x, int S, int tiles, int pitch, int c_in,
                                 float* ref) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 128 * c_in * 9) return;
    const int tap = idx % 9, ci = (idx / 9) % c_in, co = idx / (9 * c_in);
    const int shift = (tap / 3 - 1) * pitch + (tap % 3 - 1);
    const __nv_bfloat16* a = dy + (size_t(co >> 3) * S) * 8 + (co & 7);
    const __nv_bfloat16* b = x + (size_t(ci >> 3) * S) * 8 + (ci & 7);
    double acc = 0.0;
    for (int t = 0; t < tiles; ++t)
        for (int s = 0; s < 256; ++s) {
            const int s2 = s + shift;
            if (s2 < 0 || s2 >= 256) continue;
            acc += double(__bfloat162float(a[(size_t(t) * 256 + s) * 8])) * double(__bfloat162float(b[(size_t(t) * 256 + s2) * 8]));
        }
    ref[idx] = float(acc);
}

int main(int argc, char** argv) {
    const int tiles = argc > 1 ? atoi(argv[1]) : 300;
    const int c_in = argc > 2 ? atoi(argv[2]) : 128;
    const int pitch = argc > 3 ? atoi(argv[3]) : 42;
    const int S = tiles * 256;
    int dev_sms = 0;
    CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0));
    std::vector<__nv_bfloat16> h_dy(size_t(16) * S * 8), h_x(size_t(16) * S * 8);
    for (auto& v : h_dy) v = __float2bfloat16(urand() - 0.5f);
    for (size_t i = 0; i < h_x.size(); ++i) {
        const int ch = int(i / (size_t(S) * 8)) * 8 + int(i % 8);
        h_x[i] = __float2bfloat16(ch < c_in ? urand() - 0.5f : 0.f);
    }
    __nv_bfloat16 *d_dy, *d_x;
    float *d_scratch, *d_grad, *d_ref;
    const size_t wn = size_t(128) * c_in * 9;
    CK(cudaMalloc(&d_dy, h_dy.size() * 2));
    CK(cudaMalloc(&d_x, h_x.size() * 2));
    CK(cudaMalloc(&d_scratch, wgrad_scratch_elems(dev_sms) * 4));
    CK(cudaMalloc(&d_grad, wn * 4));
    CK(cudaMalloc(&d_ref, wn * 4));
    CK(cudaMemcpy(d_dy, h_dy.data(), h_dy.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_x, h_x.data(), h_x.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_grad, 0, wn * 4));
    CK(wgrad_tc_launch(d_dy, d_x, S, tiles, pitch, c_in, d_scratch, d_grad, 0, 128, 0, dev_sms, 0));
    CK(cudaDeviceSynchronize());
    wgrad_ref_kernel<<<int((wn + 127) / 128), 128>>>(d_dy, d_x, S, tiles, pitch, c_in, d_ref);
    CK(cudaDeviceSynchronize());
    std::vector<float> g(wn), ref(wn);
    CK(cudaMemcpy(g.data(), d_grad, wn * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ref.data(), d_ref, wn * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, max_ref = 0;
    size_t worst = 0;
    for (size_t i = 0; i < wn; ++i) {
        const double e = std::fabs(double(g[i]) - double(ref[i]));
        if (e > max_err) { max_err = e; worst = i; }
        max_ref = std::fmax(max_ref, std::fabs(double(ref[i])));
    }
    printf("wgrad tiles=%d c_in=%d pitch=%d: max |err| %.4g (max |ref| %.4g) at co=%zu ci=%zu tap=%zu got %.5f want %.5f\n",
           tiles, c_in, pitch, max_err, max_ref, worst / (9 * c_in), (worst / 9) % c_in, worst % 9, g[worst], ref[worst]);
    const bool ok = max_err <= 2e-3 * std::fmax(1.0, max_ref);
    // accumulate mode: a second launch doubles the gradient
    CK(wgrad_tc_launch(d_dy, d_x, S, tiles, pitch, c_in, d_scratch, d_grad, 0, 128, 1, dev_sms, 0));
    CK(cudaDeviceSynchronize());
    std::vector<float> g2(wn);
    CK(cudaMemcpy(g2.data(), d_grad, wn * 4, cudaMemcpyDeviceToHost));
    bool ok2 = true;
    for (size_t i = 0; i < wn; ++i) ok2 = ok2 && std::fabs(g2[i] - 2.f * g[i]) <= 1e-5f * std::fmax(1.f, std::fabs(g[i]));
    // timing
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int reps = 20;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i)
        CK(wgrad_tc_launch(d_dy, d_x, S, tiles, pitch, c_in, d_scratch, d_grad, 0, 128, 0, dev_sms, 0));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double flop = 2.0 * tiles * 256.0 * 128.0 * c_in * 9.0;
    printf("timing: %.1f us per wgrad (+reduce), %.1f TFLOP/s issued\n", 1e3 * ms / reps, flop / (ms / reps * 1e-3) / 1e12);
    printf(ok && ok2 ? "WGRAD SELFTEST PASSED\n" : "WGRAD SELFTEST FAILED\n");
    return ok && ok2 ? 0 : 1;
}
