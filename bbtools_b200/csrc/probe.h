// probe.h -- launchers of the per-read kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bbduk_b200.h"
#include "params.h"

// every mode, one thread per unit (pair or single read); d_units = optional list of unit indices
// d_n_units (optional): device-side count of the entries of d_units; n_units is then only an upper bound
int launch_generic(const uint8_t *d_bases, const uint32_t *d_offsets, int64_t n_units, int paired, const int32_t *d_units,
                   const unsigned int *d_n_units, const BBParams &p, const BBTable &t, const bbduk_out &out,
                   bbduk_stats *d_stats, unsigned long long *scaf_reads, unsigned long long *scaf_bases, int sm_count,
                   cudaStream_t st);

// tuned kernel (probe_fast.cu). Returns the number of kernels launched, <0 on error. Units it cannot
// handle are appended to d_handoff (count in d_handoff_n) for launch_generic.
struct FastPlan {
    bool usable;        // configuration covered by the fast kernel
    int max_read_len;   // longest read the staging is sized for
    int smem_bytes;     // dynamic shared memory per block
    int filter_words;   // words of the on-chip filter image actually used
};
FastPlan plan_fast(const BBParams &p, const BBTable &t, int max_read_len);
int launch_fast(const FastPlan &plan, const uint8_t *d_bases, const uint32_t *d_offsets, int64_t n_reads, int paired,
                const BBParams &p, const BBTable &t, const bbduk_out &out, bbduk_stats *d_stats,
                unsigned long long *scaf_reads, unsigned long long *scaf_bases, int32_t *d_handoff,
                unsigned int *d_handoff_n, int sm_count, cudaStream_t st, const uint32_t *pk_F = nullptr,
                const uint16_t *pk_D = nullptr);
// round-2 kernel (probe_fast2.cu): sampled pigeonhole scan; same contract as plan_fast / launch_fast, preferred where usable
FastPlan plan_fast2(const BBParams &p, const BBTable &t, int max_read_len);
int launch_fast2(const FastPlan &plan, const uint8_t *d_bases, const uint32_t *d_offsets, int64_t n_reads, int paired,
                 const BBParams &p, const BBTable &t, const bbduk_out &out, bbduk_stats *d_stats,
                 unsigned long long *scaf_reads, unsigned long long *scaf_bases, int32_t *d_handoff,
                 unsigned int *d_handoff_n, int sm_count, cudaStream_t st, const uint32_t *pk_F = nullptr,
                 const uint16_t *pk_D = nullptr);
// host-packed input (hostpack.h): pk_F / pk_D replace d_bases when packed_ok() and no tile can be handed off
bool packed_ok(const BBParams &p, const BBTable &t);
constexpr int FAST_MAX_READ_LEN = 1008;

// HBM-resident tables (probe_direct.cu): flat position-parallel scan -> per-read first/last hit, then
// launch_epilogue (probe_fast.cu stage D). launch_* return the number of kernels launched, <0 on error.
bool plan_direct(const BBParams &p, const BBTable &t);
size_t direct_sbits_bytes(int64_t n_bases);
int launch_direct(const uint8_t *d_bases, const uint32_t *d_offsets, int64_t n_reads, int64_t n_bases, int paired,
                  const BBParams &p, const BBTable &t, unsigned long long *d_first64, int *d_lastpos, uint16_t *d_sbits,
                  int sm_count, cudaStream_t st);
int launch_epilogue(const uint8_t *d_bases, const uint32_t *d_offsets, int64_t n_reads, int paired, const BBParams &p,
                    const BBTable &t, const bbduk_out &out, bbduk_stats *d_stats, unsigned long long *scaf_reads,
                    unsigned long long *scaf_bases, const unsigned long long *d_first64, const int *d_lastpos, int sm_count,
                    cudaStream_t st);

// trim by overlap (tbo.cu); 0 ok, 1 CUDA failure, 2 read too long
int launch_tbo(int device, int sm_count, const bbduk_tbo_cfg *cfg, const uint8_t *d_bases, const uint8_t *d_quals,
               const uint32_t *d_offsets, int64_t n_reads, int max_len, const int32_t *d_lo, int32_t *d_hi, uint8_t *d_flags,
               int32_t *d_insert, unsigned long long *d_stats, cudaStream_t st);

// quality trimming + quality / length / N filters (qtrim.cu); 0 ok, 1 CUDA failure
int launch_qtrim(int sm_count, const bbduk_qtrim_cfg *cfg, const BBParams &bp, const uint8_t *d_bases, const uint8_t *d_quals,
                 const uint32_t *d_offsets, int64_t n_reads, int paired, int32_t *d_lo, int32_t *d_hi, uint8_t *d_flags,
                 unsigned long long *d_stats, cudaStream_t st);

// low-entropy read filter (entropy.cu); 0 ok, 1 CUDA failure, 2 k / window outside the device path
int launch_entropy(int sm_count, const bbduk_entropy_cfg *cfg, const BBParams &bp, const uint8_t *d_bases, const uint32_t *d_offsets,
                   int64_t n_reads, int paired, const int32_t *d_lo, int32_t *d_hi, uint8_t *d_flags, unsigned long long *d_stats,
                   cudaStream_t st);
int launch_entropy_mask(int sm_count, const bbduk_entropy_cfg *cfg, const BBParams &bp, const uint8_t *d_bases,
                        const uint32_t *d_offsets, int64_t n_reads, int paired, int32_t *d_lo, int32_t *d_hi, uint8_t *d_flags,
                        int mode, uint32_t *d_maskbits, const int64_t *d_mask_off, unsigned long long *d_stats, cudaStream_t st);
