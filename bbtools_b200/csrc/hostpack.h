// hostpack.h -- host-side 2-bit packing of ASCII bases for the PCIe leg of bbduk_b200_process.
#pragma once
#include <stdint.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

// ASCII bases[0..n) -> F[i] = big-endian 2-bit codes of bases 16i..16i+15 (base 16i in the top bits; A0 C1 G2 T/U3,
// anything else 0) and D[i] = defined bits (bit 15-b = base 16i+b is one of ACGTUacgtu; positions >= n are 0).
// The same streams probe_fast.cu's stage A builds on the device (dna/AminoAcid.java:269-285, :1289-1320).
void pack_bases(const uint8_t *bases, int64_t n, uint32_t *F, uint16_t *D);
// groups [g0, g1) only (one worker's share)
void pack_bases_range(const uint8_t *bases, int64_t n, int64_t g0, int64_t g1, uint32_t *F, uint16_t *D);
// the same, and the groups with an undefined base listed as by list_undefined_groups (while their bits are in a register)
int64_t pack_bases_range_listing(const uint8_t *bases, int64_t n, int64_t g0, int64_t g1, uint32_t *F, uint16_t *D, uint64_t *exc,
                                 int64_t cap);
// the same as pack_bases_range, and: is every byte of the range one of A C G T N (upper case)? Then the device can spell the bases out again
// from F + D (undefined -> 'N') and nothing is lost for the steps that read ASCII (tbo, quality trimming, entropy).
bool pack_bases_range_plain(const uint8_t *bases, int64_t n, int64_t g0, int64_t g1, uint32_t *F, uint16_t *D);

// The defined bits are almost all ones: list the groups of [g0, g1) whose D is not 0xFFFF as (group << 16 | D) words.
// Returns their number, or -1 if there are more than cap (then the array itself has to travel).
int64_t list_undefined_groups(const uint16_t *D, int64_t g0, int64_t g1, uint64_t *out, int64_t cap);

// small persistent worker pool (the caller's thread takes part)
class HostPool {
  public:
    explicit HostPool(int n_threads);
    ~HostPool();
    int size() const { return (int)workers.size() + 1; }
    // runs fn(part, n_parts) on every thread of the pool and returns when all are done
    void run(const std::function<void(int, int)> &fn);

  private:
    void loop(int idx);
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    const std::function<void(int, int)> *job = nullptr;
    uint64_t epoch = 0;
    int pending = 0;
    bool stop = false;
};
