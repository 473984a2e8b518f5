// synth.cu -- device-side generator of the cfg-2 / cfg-4 synthetic read pairs.
// Same counter-based formulas as bbtools_b200/synth.py (paired_adapter_reads); tests compare the two
// byte for byte. Lets bench.py fill HBM-resident batches without PCIe traffic.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "../../include/bbduk_b200.h"

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t rnd(uint64_t seed, uint64_t stream, uint64_t idx) {
    return mix64(seed + stream * 0x9E3779B97F4A7C15ull + idx * 0xD1B54A32D192ED03ull);
}
__constant__ char kAd1[34] = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCA";
__constant__ char kAd2[34] = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT";
__constant__ char kACGT[5] = "ACGT";

__device__ __forceinline__ uint8_t with_errors(uint8_t b, uint64_t e, int sub_per_10k, int n_per_10k) {
    if ((int)((e >> 32) % 10000ull) < n_per_10k) return 'N';
    if ((int)(e % 10000ull) < sub_per_10k) {
        const uint32_t code = ((b >> 1) ^ (b >> 2)) & 3u;
        return kACGT[(code + 1u + (uint32_t)((e >> 16) % 3ull)) & 3u];
    }
    return b;
}
__device__ __forceinline__ uint8_t comp_ascii(uint8_t b) {
    return b == 'A' ? 'T' : b == 'C' ? 'G' : b == 'G' ? 'C' : b == 'T' ? 'A' : b;
}

__global__ void synth_pairs_kernel(uint8_t *bases, uint32_t *offsets, int64_t n_pairs, int64_t first_pair, int L,
                                   uint64_t seed, int sub_per_10k, int n_per_10k) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = n_pairs * L;
    if (t <= 2 * n_pairs) offsets[t] = (uint32_t)(t * L);
    if (t >= total) return;
    const int64_t lp = t / L;
    const int j = (int)(t - lp * L);
    const uint64_t p = (uint64_t)(first_pair + lp);
    const uint64_t r = rnd(seed, 0, p);
    const uint64_t cls = r % 100ull, v = r >> 8;
    const int ins = cls < 70 ? 300 + (int)(v % 301ull) : cls < 95 ? 35 + (int)(v % 115ull) : (int)(v % 35ull);
    const uint64_t g = p * (uint64_t)L + (uint64_t)j;
    uint8_t b1 = kACGT[rnd(seed, 1, g) & 3ull];
    uint8_t b2 = kACGT[rnd(seed, 2, g) & 3ull];
    if (ins < L) {
        if (j < ins) b2 = comp_ascii((uint8_t)kACGT[rnd(seed, 1, p * (uint64_t)L + (uint64_t)(ins - 1 - j)) & 3ull]);
        const int ai = j - ins;
        if (ai >= 0 && ai < 33) {
            b1 = kAd1[ai];
            b2 = kAd2[ai];
        }
    }
    b1 = with_errors(b1, rnd(seed, 3, g), sub_per_10k, n_per_10k);
    b2 = with_errors(b2, rnd(seed, 4, g), sub_per_10k, n_per_10k);
    bases[(2 * lp) * L + j] = b1;
    bases[(2 * lp + 1) * L + j] = b2;
}

// cfg 3 / cfg 4 reference: base i = ACGT[rnd(seed, 5, i) & 3]  (synth.py:random_reference)
__global__ void synth_reference_kernel(uint8_t *out, int64_t n, uint64_t seed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = kACGT[rnd(seed, 5, (uint64_t)i) & 3ull];
}

// cfg 3 reads (synth.py:contaminant_reads): contam_pct % copied from the reference (forward strand, uniform
// start over the concatenated reference), the rest uniform ACGT; then substitutions and N's
__global__ void synth_contam_kernel(uint8_t *bases, uint32_t *offsets, int64_t n_reads, int64_t first_read, int L,
                                    const uint8_t *__restrict__ ref, int64_t ref_len, uint64_t seed, int contam_pct,
                                    int sub_per_10k, int n_per_10k) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= n_reads) offsets[t] = (uint32_t)(t * L);
    if (t >= n_reads * L) return;
    const int64_t lr = t / L;
    const int j = (int)(t - lr * L);
    const uint64_t r = (uint64_t)(first_read + lr);
    const uint64_t g = r * (uint64_t)L + (uint64_t)j;
    uint8_t b = kACGT[rnd(seed, 1, g) & 3ull];
    const uint64_t c = rnd(seed, 0, r);
    if ((int)(c % 100ull) < contam_pct) {
        const uint64_t start = (c >> 8) % (uint64_t)(ref_len - L + 1);
        b = ref[start + (uint64_t)j];
    }
    bases[t] = with_errors(b, rnd(seed, 3, g), sub_per_10k, n_per_10k);
}

}  // namespace

extern "C" BBDUK_API int bbduk_b200_synth_reference(uint8_t *d_out, int64_t n, uint64_t seed, void *stream) {
    if (n <= 0) return 1;
    synth_reference_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_out, n, seed);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

extern "C" BBDUK_API int bbduk_b200_synth_contam(uint8_t *d_bases, uint32_t *d_offsets, int64_t n_reads, int64_t first_read,
                                                 int32_t read_len, const uint8_t *d_ref, int64_t ref_len, uint64_t seed,
                                                 int32_t contam_pct, int32_t sub_per_10k, int32_t n_per_10k, void *stream) {
    if (n_reads <= 0 || read_len <= 0 || ref_len < read_len || !d_ref) return 1;
    if (n_reads * (int64_t)read_len >= (1ll << 32)) return 1;
    const int64_t threads = std::max<int64_t>(n_reads * read_len, n_reads + 1);
    synth_contam_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_bases, d_offsets, n_reads, first_read, read_len, d_ref, ref_len, seed, contam_pct, sub_per_10k, n_per_10k);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

extern "C" BBDUK_API int bbduk_b200_synth_pairs(uint8_t *d_bases, uint32_t *d_offsets, int64_t n_pairs, int64_t first_pair,
                                                int32_t read_len, uint64_t seed, int32_t sub_per_10k, int32_t n_per_10k,
                                                void *stream) {
    if (n_pairs <= 0 || read_len <= 0) return 1;
    if (2 * n_pairs * (int64_t)read_len >= (1ll << 32)) return 1;
    const int64_t threads = std::max<int64_t>(n_pairs * read_len, 2 * n_pairs + 1);
    synth_pairs_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_bases, d_offsets, n_pairs, first_pair, read_len, seed, sub_per_10k, n_per_10k);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
