// table.h -- host-side owner of the device hash array (see table.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "params.h"

struct DeviceTable {
    uint64_t *d_keys = nullptr;
    int32_t *d_vals = nullptr;
    uint32_t *d_filter = nullptr;
    int64_t n_slots = 0;
    uint32_t n_filter_words = 0;  // canonical bloom words; the buffer also holds part_words + short_words
    uint32_t part_words = 0, short_words = 0;
    uint32_t samp_words = 0, tail_words = 0;  // probe_fast2.cu: 8-mer byte map, tail bitmaps (params.h)
    int32_t tail_q = 0;
    uint32_t big_words = 0;  // L2-resident filter for HBM-resident arrays, last segment of d_filter
    int32_t n_parts = 0, part_w = 0, part_lag[4] = {0, 0, 0, 0};
    uint32_t total_filter_words() const { return n_filter_words + part_words + short_words + samp_words + tail_words + big_words; }
    int32_t n_scaffolds = 0;
    int64_t stored = 0;     // distinct keys ("Added N kmers", jgi/BBDuk.java:1973)
    int64_t ref_kmers = 0;  // refKmers (jgi/BBDuk.java:1956)
    bool owns = true;

    int alloc(int64_t slots, uint32_t total_filter_words, char *err, int errlen);
    void release();
    BBTable view() const;
    // ref = concatenated scaffolds (ids 1..n in order)
    int build(const BBParams &p, const std::vector<uint8_t> &ref, const std::vector<int64_t> &offsets, int load_pct,
              uint32_t filter_words, cudaStream_t st, int64_t *launches, char *err, int errlen);
};
