// probe_fast.cu -- placeholder until the tuned kernel lands: reports "not usable" so every batch takes
// the generic kernel.
#include "probe.h"

FastPlan plan_fast(const BBParams &, const BBTable &, int) { return FastPlan{false, 0, 0, 0}; }
int launch_fast(const FastPlan &, const uint8_t *, const uint32_t *, int64_t, int, const BBParams &, const BBTable &,
                const bbduk_out &, bbduk_stats *, unsigned long long *, unsigned long long *, int32_t *, unsigned int *,
                int, cudaStream_t) {
    return -1;
}
