// sort.cuh -- internal interface of the onesweep radix sort and the chained scan (sort.cu)
#pragma once
#include "common.cuh"

// Stable LSD radix sort of (key, value) pairs on `key_bits` low key bits, 8-bit digits, one histogram kernel for all
// passes plus one onesweep kernel per pass.  `count` is a device word, max_count the host-known upper bound.
// values_in == nullptr means "values are 0..count-1" (write_sequence fused away).
// Never writes to keys_in / values_in.  Result in keys_out / values_out.
int apbf_radix_sort_pairs(apbf_ctx* ctx, const uint32_t* keys_in, const uint32_t* values_in, uint32_t* keys_out,
                          uint32_t* values_out, const uint32_t* count, uint32_t max_count, int key_bits);

// number of key bits the reference's 4-bit LSD passes would look at for this upper bound (algorithms.cpp:73)
int apbf_reference_sort_bits(uint32_t upper_bound);

// result[i] = sum_{j<=i} values[j] (inclusive) or sum_{j<i} (exclusive); optionally writes the grand total
// to *total_out (device) -- clamped to total_clamp -- and raises flag bit 0 in *flags_out when it was clamped.
int apbf_scan_u32(apbf_ctx* ctx, const uint32_t* values, uint32_t* result, const uint32_t* count, uint32_t max_count,
                  bool inclusive, uint32_t* total_out, uint32_t total_clamp, uint32_t* flags_out, uint32_t* raw_total_out,
                  const uint32_t* run_if = nullptr); // run_if: device word; the scan does nothing when it is 0
