// Exercises include/sylow_b200.hpp (the C++ host mirror) on the GPU; prints values the pytest wrapper
// compares with the golden vectors / the oracle.  Reads like the reference's own tests
// (src/pairing.rs:1052-1072,1101-1120).
#include <cstdio>

#include "../../include/sylow_b200.hpp"

using namespace sylow;

static void print_fp12(const char* tag, const Fp12& f) {
  std::printf("%s", tag);
  for (const Fp& c : f.c) std::printf(" %016llx%016llx%016llx%016llx", (unsigned long long)c.w[3], (unsigned long long)c.w[2],
                                      (unsigned long long)c.w[1], (unsigned long long)c.w[0]);
  std::printf("\n");
}

int main() {
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
    return 0;
  } catch (const Error& e) {
    std::printf("ERROR %d %s\n", e.status, e.what());
    return 1;
  }
}
