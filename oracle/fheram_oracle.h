/*
 * fheram_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of phantomzone-org/fhe-ram's homomorphic read /
 * read_prepare_write / write path (src/ram.rs, src/coordinate*.rs, src/address.rs,
 * src/base.rs, src/keys.rs, src/parameters.rs, examples/fhe-ram.rs) plus the Poulpy
 * 0.3.2 layer those files call (poulpy-core / poulpy-hal / poulpy-backend, NOT present
 * in /root/reference: un-vendored path dependency, Cargo.toml:7-10, Cargo.lock:376-434).
 *
 * PARITY UNPINNED at the ciphertext-limb level: the reference ships no golden
 * ciphertexts and its arithmetic dependency cannot be built here, so the Poulpy layer is
 * restated from its published algorithm (see oracle/SPEC.md for every convention that was
 * chosen).  What IS pinned: the reference's own KATs for the digit layout
 * (src/base.rs:114-438) and parameters (src/parameters.rs:296-323), and the acceptance
 * scenario of examples/fhe-ram.rs:104-115,127-138,165-176 (decrypt == plaintext, noise
 * bound) -- see tests/test_oracle_*.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this library.  The product (fhe_ram_b200/) never links or calls it.
 *
 * Two builds of the same source:
 *   liboracle_exact.so  negacyclic products by a 62-bit-prime NTT (exact integers)
 *   liboracle_fft64.so  negacyclic products by an f64 FFT (the reference's FFT64 scheme);
 *                       used as the timed CPU baseline; cross-checked against _exact.
 *
 * Data layouts (all int64, identical to the product's C ABI):
 *   VecZnx(n, cols, size): index ((limb*cols)+col)*n + coeff, limb 0 most significant.
 *   GLWE(k): cols = 2 (col 0 body, col 1 mask), size = ceil(k/base2k).
 *   GGSW: [dnum rows][2 cols_in] GLWE(k_addr).   GGLWE key: [dnum rows] GLWE(k_key).
 */
#ifndef FHERAM_ORACLE_H
#define FHERAM_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_params {
  int32_t log_n;          /* src/parameters.rs:11 LOG_N */
  int32_t base2k;         /* :12 */
  int32_t k_pt;           /* :14 K_GLWE_PT (snapshot 3, README 9) */
  int32_t k_ct;           /* :15 */
  int32_t k_addr;         /* :16 K_GGSW_ADDR */
  int32_t k_evk_trace;    /* :17 */
  int32_t k_evk_ggsw_inv; /* :18 */
  int32_t word_size;      /* :20 */
  int32_t n_decomp;       /* :19 DECOMP_N */
  int32_t decomp_n[8];
  uint64_t max_addr;      /* :21 */
} orc_params;

void orc_params_snapshot(orc_params *p); /* src/parameters.rs:11-21 */
void orc_params_readme(orc_params *p);   /* README.md:17-34 */

/* src/base.rs:84-108 get_base_2d; returns number of coordinates; lens[c] digits each */
int orc_get_base_2d(uint32_t value, const int32_t *base, int n_base, int32_t *lens,
                    int32_t *digits /* [8][8] flat */);
uint32_t orc_base1d_max(const int32_t *b, int n);                       /* base.rs:10-14 */
uint32_t orc_base1d_gap(const int32_t *b, int n, int log_n);             /* base.rs:17-21 */
void orc_base1d_decomp(const int32_t *b, int n, uint32_t v, uint8_t *o); /* base.rs:24-33 */
uint32_t orc_base1d_recomp(const int32_t *b, int n, const uint8_t *d);   /* base.rs:36-44 */
uint64_t orc_reverse_bits_msb(uint64_t x, uint32_t n);                   /* src/lib.rs:23-26 */

typedef struct orc_ctx orc_ctx;
typedef struct orc_source orc_source;
typedef struct orc_keys orc_keys;
typedef struct orc_ram orc_ram;
typedef struct orc_packer orc_packer;

orc_ctx *orc_ctx_new(const orc_params *p);
void orc_ctx_free(orc_ctx *c);
const char *orc_backend_name(void);

/* derived sizes, in int64 words */
size_t orc_n(const orc_ctx *c);
size_t orc_glwe_len(const orc_ctx *c);     /* GLWE(k_ct) */
size_t orc_ggsw_len(const orc_ctx *c);     /* GGSW(k_addr, dnum_ct) */
size_t orc_atk_len(const orc_ctx *c);      /* trace automorphism key */
size_t orc_evk_inv_len(const orc_ctx *c);  /* atk_ggsw_inv / tsk_ggsw_inv */
int orc_n_gal(const orc_ctx *c);           /* log_n trace keys */
int orc_n_ggsw(const orc_ctx *c);          /* GGSWs per address */
int orc_n_glwe_per_subram(const orc_ctx *c);
int64_t orc_gal_el(const orc_ctx *c, int i); /* GLWE::trace_galois_elements[i] */

/* PRNG (stands in for poulpy_hal::source::Source; ChaCha20 keystream) */
orc_source *orc_source_new(const uint8_t seed[32]);
void orc_source_free(orc_source *s);
uint64_t orc_source_next_u64(orc_source *s);
uint32_t orc_source_next_u32(orc_source *s);
void orc_source_fill_bytes(orc_source *s, uint8_t *out, size_t n);

/* ---- client side (examples/fhe-ram.rs:34-95, 179-237) ---- */
void orc_secret_gen(const orc_ctx *c, orc_source *xs, int64_t *sk /* n */);
/* src/keys.rs:135-180 */
void orc_keygen(const orc_ctx *c, const int64_t *sk, orc_source *xa, orc_source *xe,
                int64_t *atk_glwe /* n_gal*atk_len */, int64_t *tsk /* evk_inv_len */,
                int64_t *atk_inv /* evk_inv_len */);
/* src/ram.rs:129-167,334-380 ; out = [word_size][n_glwe_per_subram] GLWE */
void orc_ram_encrypt(const orc_ctx *c, const uint8_t *data, const int64_t *sk, orc_source *xa,
                     orc_source *xe, int64_t *out);
/* src/address.rs:86-109 + src/coordinate.rs:121-180 ; out = n_ggsw GGSW */
void orc_address_encrypt(const orc_ctx *c, uint32_t value, const int64_t *sk, orc_source *xa,
                         orc_source *xe, int64_t *out);
/* examples/fhe-ram.rs:179-210 encrypt_glwe */
void orc_encrypt_byte(const orc_ctx *c, uint8_t value, const int64_t *sk, orc_source *xa,
                      orc_source *xe, int64_t *out);
/* examples/fhe-ram.rs:212-237 decrypt_glwe */
void orc_decrypt_glwe(const orc_ctx *c, const int64_t *glwe, const int64_t *sk, int64_t want,
                      int64_t *value, double *noise);
/* full plaintext (size_ct limbs x n), b + a*s normalized */
void orc_glwe_decrypt(const orc_ctx *c, const int64_t *glwe, const int64_t *sk, int64_t *pt);
/* GGSW noise helper: decrypts row r / col_in ci of a GGSW to a plaintext (size_addr limbs) */
void orc_ggsw_decrypt_row(const orc_ctx *c, const int64_t *ggsw, int row, int col_in,
                          const int64_t *sk, int64_t *pt);
int64_t orc_cast_u8_to_signed(uint8_t v, int bits); /* examples/fhe-ram.rs:25-32 */

/* ---- evaluation keys (src/keys.rs:27-71) ---- */
orc_keys *orc_keys_prepare(const orc_ctx *c, const int64_t *atk_glwe, const int64_t *tsk,
                           const int64_t *atk_inv);
void orc_keys_free(orc_keys *k);

/* ---- Poulpy-level ops, exposed for kernel parity tests ---- */
void orc_glwe_normalize(const orc_ctx *c, int64_t *glwe);
void orc_glwe_rsh(const orc_ctx *c, int k, int64_t *glwe);
void orc_glwe_rotate(const orc_ctx *c, int64_t k, const int64_t *in, int64_t *out);
void orc_glwe_small_automorphism(const orc_ctx *c, int64_t p, const int64_t *in, int64_t *out);
void orc_external_product(const orc_ctx *c, const int64_t *in, const int64_t *ggsw, int64_t *out);
/* chain of n external products (CoordinatePrepared::product, coordinate_prepared.rs:147-161) */
void orc_coordinate_product(const orc_ctx *c, const int64_t *in, const int64_t *ggsws, int n,
                            int64_t *out);
/* mode 0: glwe_automorphism, 1: automorphism_add (res = a + phi(a)),
 *      2: automorphism_sub_negate (res = a - phi(a)); gal index i (key for orc_gal_el(i)) */
void orc_automorphism(const orc_ctx *c, const orc_keys *k, int gal_idx, int mode,
                      const int64_t *in, int64_t *out);
void orc_trace(const orc_ctx *c, const orc_keys *k, int start, int end, const int64_t *in,
               int64_t *out);
orc_packer *orc_packer_new(const orc_ctx *c);
void orc_packer_free(orc_packer *p);
void orc_packer_add(orc_packer *p, const orc_keys *k, const int64_t *glwe_or_null);
void orc_packer_flush(orc_packer *p, int64_t *out);
/* GLWEPacker::combine at tree level `level`: a (holds a value) updated in place, b may be NULL */
void orc_packer_combine(const orc_ctx *c, const orc_keys *k, int level, int64_t *a, const int64_t *b);
/* GGSW(X^i) -> GGSW(X^-i) (coordinate_prepared.rs:121-142, raw output before prepare) */
void orc_ggsw_automorphism_inv(const orc_ctx *c, const orc_keys *k, const int64_t *in,
                               int64_t *out);

/* ---- Ram (src/ram.rs) ---- */
orc_ram *orc_ram_new(const orc_ctx *c);
void orc_ram_free(orc_ram *r);
void orc_ram_load(orc_ram *r, const int64_t *cts);   /* Ram::encrypt_sk result */
void orc_ram_store(const orc_ram *r, int64_t *cts);
void orc_ram_tree_store(const orc_ram *r, int64_t *cts /* [word_size] GLWE, tree.last()[0] */);
int orc_ram_state(const orc_ram *r);
/* return 0 ok, <0 = the reference's assert would panic */
void orc_external_product_many(const orc_ctx *c, const int64_t *in, int n, const int64_t *ggsw, int64_t *out, int threads);
void orc_set_ram_threads(int n); /* sub-RAMs of one call on n threads (checker speed only; default 1) */
int orc_ram_read(orc_ram *r, const int64_t *addr, const orc_keys *k, int64_t *out);
int orc_ram_read_prepare_write(orc_ram *r, const int64_t *addr, const orc_keys *k, int64_t *out);
int orc_ram_write(orc_ram *r, const int64_t *w, const int64_t *addr, const orc_keys *k);
/* n independent reads run on `threads` OpenMP threads (CPU-baseline helper; each read is the
 * sequential reference algorithm on a private copy of the packer state) */
int orc_ram_read_many(orc_ram *r, const int64_t *addrs, int n, const orc_keys *k, int64_t *out,
                      int threads);
/* counters of vmp-class ops executed since ctx creation: [0]=external products, [1]=key-switches */
void orc_op_counters(const orc_ctx *c, uint64_t out[2]);

#ifdef __cplusplus
}
#endif
#endif
