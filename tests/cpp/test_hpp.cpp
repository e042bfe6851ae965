// Exercises include/sylow_b200.hpp (the C++ host mirror) on the GPU; prints values the pytest wrapper
// compares with the golden vectors / the oracle.  Reads like the reference's own tests
// (src/pairing.rs:1052-1072,1101-1120).
#include <cstdio>
#include <cstdlib>

#include "../../include/sylow_b200.hpp"

using namespace sylow;

static void print_fp12(const char* tag, const Fp12& f) {
  std::printf("%s", tag);
  for (const Fp& c : f.c) std::printf(" %016llx%016llx%016llx%016llx", (unsigned long long)c.w[3], (unsigned long long)c.w[2],
                                      (unsigned long long)c.w[1], (unsigned long long)c.w[0]);
  std::printf("\n");
}

int main(int argc, char** argv) {
  const int n_gpus = argc > 1 ? std::atoi(argv[1]) : 1;
  try {
    Engine eng(0);
    G1Affine g1{Fp::from_u64(1), Fp::from_u64(2), false};
    G2Affine g2;  // src/groups/g2.rs:47-77
    g2.x.c0.w = {5106727233969649389ull, 7440829307424791261ull, 4785637993704342649ull, 1729627375292849782ull};
    g2.x.c1.w = {10945020018377822914ull, 17413811393473931026ull, 8241798111626485029ull, 1841571559660931130ull};
    g2.y.c0.w = {5541340697920699818ull, 16416156555105522555ull, 5380518976772849807ull, 1353435754470862315ull};
    g2.y.c1.w = {6173549831154472795ull, 13567992399387660019ull, 17050234209342075797ull, 650358724130500725ull};
    // test_gt_generator
    Gt gt = eng.pairing(g1, g2);
    print_fp12("GT", gt.v);
    // test_identities
    G1Affine inf1{Fp::from_u64(0), Fp::from_u64(1), true};
    std::printf("IDENT %d\n", eng.pairing(inf1, g2) == Gt::identity());
    std::printf("EMPTY %d\n", eng.glued_pairing({}, {}) == Gt::identity());
    // test_signing: sign -> verify, wrong message fails
    Fp sk = Fp::from_u64(0x1234567890abcdefull);
    std::string msg("\x00\x00\x00\x14", 4);
    G1Affine sig = eng.sign_batch({sk}, {msg})[0];
    G2Affine pk = eng.g2_mul_batch({g2}, {sk})[0];
    std::printf("SIG %016llx%016llx%016llx%016llx\n", (unsigned long long)sig.x.w[3], (unsigned long long)sig.x.w[2],
                (unsigned long long)sig.x.w[1], (unsigned long long)sig.x.w[0]);
    std::printf("VERIFY %d %d %d\n", (int)eng.verify(pk, msg, sig), (int)eng.verify(pk, "other", sig),
                (int)eng.verify_batch({pk, pk}, {msg, msg}, {sig, sig}));
    // bilinearity: e(kP, Q) == e(P, kQ)
    Fp k = Fp::from_u64(987654321);
    Gt a = eng.pairing(eng.g1_mul_batch({g1}, {k})[0], g2), b = eng.pairing(g1, eng.g2_mul_batch({g2}, {k})[0]);
    std::printf("BILINEAR %d\n", a == b && !(a == Gt::identity()));
    // MultiEngine (sylow_b200_create_multi): two device slots must agree with one - contiguous slices, concatenated
    // results, and the 384-byte partial exchange of verify_batch.  Slots {0, 0} share GPU 0; with a second GPU visible
    // the same checks run on {0, 1}.
    for (int leg = 0; leg < (n_gpus >= 2 ? 2 : 1); leg++) {
      MultiEngine multi(leg == 0 ? std::vector<int>{0, 0} : std::vector<int>{0, 1});
      std::vector<G1Affine> ps;
      std::vector<G2Affine> qs;
      std::vector<Fp> ks;
      for (int i = 0; i < 5; i++) ks.push_back(Fp::from_u64(1000003ull * (i + 1)));
      ps = eng.g1_mul_batch(std::vector<G1Affine>(5, g1), ks);
      qs = eng.g2_mul_batch(std::vector<G2Affine>(5, g2), ks);
      std::vector<Gt> one = eng.pairing_batch(ps, qs), two = multi.pairing_batch(ps, qs);
      bool same_gt = one.size() == two.size();
      for (size_t i = 0; i < one.size() && same_gt; i++) same_gt = one[i] == two[i];
      std::vector<G1Affine> m1 = multi.g1_mul_batch(std::vector<G1Affine>(5, g1), ks);
      bool same_mul = true;
      for (size_t i = 0; i < 5; i++) same_mul = same_mul && m1[i].x == ps[i].x && m1[i].y == ps[i].y;
      std::vector<std::string> msgs = {"a", "bb", "ccc", "dddd", "eeeee"};
      std::vector<G1Affine> sigs = eng.sign_batch(ks, msgs);
      std::vector<G2Affine> pks = eng.g2_mul_batch(std::vector<G2Affine>(5, g2), ks);
      bool good = multi.verify_batch(pks, msgs, sigs);
      std::swap(sigs[1], sigs[3]);
      sigs[1] = sigs[0];
      bool bad = multi.verify_batch(pks, msgs, sigs);
      // random-weight batch verification: same verdicts, and the cancelling forgery (sig_0 + D, sig_1 - D) that the
      // unweighted product accepts is rejected
      std::array<std::uint8_t, 32> seed{};
      for (int i = 0; i < 32; i++) seed[i] = (std::uint8_t)(17 * i + 3);
      std::vector<G1Affine> good_sigs = eng.sign_batch(ks, msgs);
      bool wgood = multi.verify_batch(pks, msgs, good_sigs, DST(), &seed);
      bool wbad = multi.verify_batch(pks, msgs, sigs, DST(), &seed);
      std::printf("MULTI%d %d %d %d %d %d %d %zu\n", leg, (int)same_gt, (int)same_mul, (int)good, (int)bad, (int)wgood,
                  (int)wbad, multi.size());
    }
    return 0;
  } catch (const Error& e) {
    std::printf("ERROR %d %s\n", e.status, e.what());
    return 1;
  }
}
