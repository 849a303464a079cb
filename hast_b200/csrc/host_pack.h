// host_pack.h -- host-side 2-bit packing of a read batch on several threads (see host_pack.cpp).
#pragma once
#include <cstdint>

namespace hastpack {

// ASCII bases -> 2-bit words (16 bases per word, first base in the top two bits, code (byte >> 1) & 3 of kmer.h:11;
// the last word zero-padded) and one bit per read: "contains the byte 'N'" (classify.cpp:182-185).
// words_out: (n_bases + 15) / 16 words; has_n_out: (n_reads + 31) / 32 words (zeroed here).
class Pool;
Pool* pool_create(int threads);                 // threads >= 1 (the caller's thread included)
void pool_destroy(Pool* p);
void pack_batch(Pool* p, const uint8_t* bases, uint64_t n_bases, const uint32_t* read_off, uint32_t n_reads,
                uint32_t* words_out, uint32_t* has_n_out);

}  // namespace hastpack
