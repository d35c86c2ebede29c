// Drives playsnark_b200/host/playsnark.hpp on the README circuit from a token file written by
// tests/test_host_cpp.py (golden fixture) and prints the proof as hex for the test to compare.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>

#include "../playsnark_b200/host/playsnark.hpp"

using namespace playsnark;

static std::ifstream in;
template <size_t N> static std::array<uint8_t, N> rd() {
  std::string s; in >> s;
  std::array<uint8_t, N> a{};
  if (s.size() != 2 * N) { fprintf(stderr, "bad token length %zu (want %zu)\n", s.size(), 2 * N); exit(2); }
  for (size_t i = 0; i < N; i++) a[i] = (uint8_t)std::stoi(s.substr(2 * i, 2), nullptr, 16);
  return a;
}
template <size_t N> static std::vector<std::array<uint8_t, N>> rdv() {
  size_t n; in >> n;
  std::vector<std::array<uint8_t, N>> v(n);
  for (auto& x : v) x = rd<N>();
  return v;
}
template <size_t N> static void pr(const char* name, const std::array<uint8_t, N>& a) {
  printf("%s ", name);
  for (auto b : a) printf("%02x", b);
  printf("\n");
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  in.open(argv[1]);
  QAP q;
  in >> q.nbVars >> q.nbIO >> q.nbGates;
  for (auto* ps : {&q.left, &q.right, &q.out}) { ps->resize(q.nbVars); for (auto& p : *ps) p = rdv<32>(); }
  q.z = rdv<32>();
  Vector sol; { size_t n; in >> n; sol.resize(n); for (auto& v : sol) in >> v; }
  Groth16Setup tr;
  tr.Alpha = rd<48>(); tr.Beta = rd<48>(); tr.Delta = rd<48>(); tr.Beta2 = rd<96>(); tr.Delta2 = rd<96>();
  tr.Xi = rdv<48>(); tr.Xi2 = rdv<96>(); tr.XiT = rdv<48>(); tr.NioLP = rdv<48>();
  Element r = rd<32>(), s = rd<32>();
  PHGR13EvalKey ek;
  ek.gsi = rdv<48>(); ek.vs = rdv<48>(); ek.ws = rdv<96>(); ek.ys = rdv<48>(); ek.vas = rdv<48>(); ek.was = rdv<48>();
  ek.yas = rdv<48>(); ek.vbs = rdv<48>(); ek.wbs = rdv<48>(); ek.ybs = rdv<48>();
  try {
    Backend be(0);
    ResidentQAP rq(be, q);
    Poly h = rq.Quotient(sol);
    for (auto& c : h) pr("h", c);
    Groth16Prover g16(be, tr);
    Groth16Proof p = g16.Prove(rq, sol, r, s);
    pr("A", p.A); pr("B", p.B); pr("C", p.C);
    PHGR13Prover ph(be, ek);
    PHGR13Proof pp = ph.Prove(rq, sol);
    pr("hs", pp.hs); pr("vss", pp.vss); pr("wss", pp.wss); pr("yss", pp.yss); pr("vass", pp.vass); pr("wass", pp.wass);
    pr("yass", pp.yass); pr("gz", pp.gz);
    BlindedPoints bp(be, tr.XiT);
    pr("htd", bp.BlindEval(h));
    // error behaviour: invalid witness -> "apocalypse"; wrong length -> length_error
    Vector bad = sol; bad[3] += 1;
    try { rq.Quotient(bad); printf("err none\n"); } catch (const std::runtime_error& e) { printf("err %s\n", e.what()); }
    Poly shortp(h.begin(), h.end() - 1);
    try { bp.BlindEval(shortp); printf("len none\n"); } catch (const std::length_error&) { printf("len mismatch\n"); }
    pr("neg", ToFieldElement(-1));
  } catch (const std::exception& e) {
    fprintf(stderr, "FAILED: %s\n", e.what());
    return 1;
  }
  return 0;
}
