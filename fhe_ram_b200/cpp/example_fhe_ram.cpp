// C++ restatement of the reference's acceptance scenario (examples/fhe-ram.rs:34-177) on the
// B200 path: keygen, encrypt RAM + address, read, read_prepare_write, write, read back, with the
// example's decrypt == plaintext and noise assertions.  Usage: example_fhe_ram [log2(max_addr)]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "fheram.hpp"

using namespace fheram;
using clk = std::chrono::steady_clock;

static int64_t cast_u8_to_signed(uint8_t v, int bits) {  // examples/fhe-ram.rs:25-32
  int shift = 8 - bits;
  return (int64_t)((int8_t)(uint8_t)(v << shift)) >> shift;
}

int main(int argc, char** argv) {
  try {
    Parameters params = Parameters::new_();
    if (argc > 1) params.c.max_addr = 1ull << atoi(argv[1]);
    params.c.k_pt = 8;
    Source xs = Source::filled(0), xa = Source::filled(0), xe = Source::filled(0);  // :37-43
    GLWESecret sk = GLWESecret::fill_ternary_prob(params, 0.5, xs);                 // :49-50
    EvaluationKeys keys = EvaluationKeys::encrypt_sk(params, sk, xa, xe);           // :52-53
    EvaluationKeysPrepared kp = EvaluationKeysPrepared::alloc(params);              // :61-63
    kp.prepare(keys);
    Source source = Source::filled(5);                                              // :66
    const size_t ws = params.word_size();
    std::vector<uint8_t> data(params.max_addr() * ws);
    source.fill_bytes(data);                                                        // :72-73
    Ram ram(params);                                                                // :76
    ram.encrypt_sk(data, sk, xa, xe);                                               // :79
    Address addr = Address::alloc_from_params(params);                              // :82
    uint32_t idx = source.next_u32() % (uint32_t)params.max_addr();                 // :85
    addr.encrypt_sk(params, idx, sk, xa, xe);                                       // :88-95

    auto check = [&](const std::vector<GLWE>& ct, const std::vector<uint8_t>& d) {  // :104-115
      for (size_t i = 0; i < ws; i++) {
        int64_t want = cast_u8_to_signed(d[i + ws * idx], params.k_glwe_pt()), v;
        double noise;
        fheram::check(fheram_decrypt_word(&params.c, ct[i].data(), sk.data.data(), want, &v, &noise));
        printf("noise: %f\n", noise);
        if (v != want || !(noise < -(params.k_glwe_pt() + 1.0))) {
          fprintf(stderr, "MISMATCH byte %zu: got %ld want %ld noise %f\n", i, (long)v, (long)want, noise);
          exit(2);
        }
      }
    };
    auto ms = [](clk::time_point a) { return std::chrono::duration<double, std::milli>(clk::now() - a).count(); };

    auto t = clk::now();
    auto ct = ram.read(addr, kp);                                                   // :98-101
    printf("READ Elapsed time: %.3f ms\n", ms(t));
    check(ct, data);
    t = clk::now();
    ct = ram.read_prepare_write(addr, kp);                                          // :118-124
    printf("READ_PREPARE_WRITE Elapsed time: %.3f ms\n", ms(t));
    check(ct, data);
    std::vector<uint8_t> value(ws);
    source.fill_bytes(value);                                                       // :141-142
    std::vector<GLWE> ct_w;
    for (auto b : value) {                                                          // :145-148, 179-210
      Source a1 = Source::filled(1), e1 = Source::filled(1);
      GLWE g(params.glwe_len());
      fheram::check(fheram_encrypt_word(&params.c, b, sk.data.data(), a1.raw(), e1.raw(), g.data()));
      ct_w.push_back(g);
    }
    t = clk::now();
    ram.write(ct_w, addr, kp);                                                      // :151-154
    printf("WRITE Elapsed time: %.3f ms\n", ms(t));
    for (size_t i = 0; i < ws; i++) data[i + ws * idx] = value[i];                  // :157-159
    ct = ram.read(addr, kp);                                                        // :162
    check(ct, data);
    printf("OK\n");
    return 0;
  } catch (const Panic& e) {
    fprintf(stderr, "panic [%d]: %s\n", e.code, e.what());
    return 1;
  }
}
