// Internal interface between the client-side code (client.cpp) and the GPU bulk-encryption path
// (fheram_cuda.cu).  Not part of the C ABI.
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/fheram.h"

// key and absolute position (in 32-bit words) of a Source's ChaCha20 stream
void fheram_source_tell(const fheram_source* s, uint32_t key[8], uint64_t* word_pos);
// advance the stream by n 32-bit words without producing them
void fheram_source_skip_words(fheram_source* s, uint64_t n);
// n draws of the GLWE encryption noise (sigma 3.2, bound 6 sigma), exactly as glwe_encrypt draws them
void fheram_source_noise_i8(fheram_source* s, int8_t* out, size_t n);
// the noise draw that starts `word_offset` words past the Source's position (the Source is not advanced)
int8_t fheram_source_noise_at(const fheram_source* s, uint64_t word_offset);
// the n_ggsw monomials +/- X^pos an address value is encoded as (src/address.rs:102-108,
// src/coordinate.rs:148-179); returns n_ggsw or a negative status
int fheram_address_monomials(const fheram_params* p, uint32_t value, int32_t* pos, int32_t* sign);
// range checks of every field the digit tables / size helpers index with (n_decomp <= 8, max_addr <= N^2, ...);
// 0 or FHERAM_ERR_INVALID with fheram_last_error set
int fheram_params_check(const fheram_params* p);
