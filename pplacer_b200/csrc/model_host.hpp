// Host-side model numerics of the placement path (C++; no CUDA here).
//   discrete Gamma rates        - Gamma.discrete_gamma          pplacer_src/gamma.ml:52-70
//   Q = B diag(pi) eigen form   - Diagd.normed_of_exchangeable_pair  pplacer_src/diagd.ml:123-172
//   empirical frequencies       - Alignment.emper_freq          common_src/alignment.ml:186-217
//   character -> state set      - nuc_models.ml:24-54, prot_models.ml:5-51
// GSL (gamma_P/Pinv, gamma_inc_P, eigen_symmv) is an un-vendored dependency of the reference; the
// routines below are independent implementations of the same mathematical functions.
#pragma once
#include <cstdint>
#include <vector>

namespace ppb {

constexpr int kMaxStates = 20;
constexpr int kMaxCodes = 32;

// regularised lower incomplete gamma P(a, x) and its inverse in x
double gamma_p(double a, double x);
double gamma_p_inv(double p, double a);
void discrete_gamma(int n_cat, double alpha, double *rates_out);

// symmetric eigen-decomposition, eigenvectors in the columns of evec (row-major n x n)
void jacobi_eigen(int n, const double *sym, double *eval, double *evec);
// exch: S x S symmetric (diagonal ignored); pi: S.  Outputs x (S x S), lambda (S), xit (S x S).
void diagd_normed(int S, const double *exch, const double *pi, double *x, double *lambda, double *xit);

// Alphabet coding used on the device: every character maps to a small code; code -> bitmask of
// compatible states.  Bit 7 of a code marks an uninformative site ('-' or '?').
struct Alphabet {
  int n_states = 0;
  int n_codes = 0;               // number of distinct informative codes
  uint8_t code_of_char[256];     // 0xFF = unknown character
  uint32_t mask_of_code[kMaxCodes];
  int any_code = 0;              // the all-ones code (N / X); gaps map to (any_code | 0x80)
};
constexpr uint8_t kGapBit = 0x80;
constexpr uint8_t kBadCode = 0xFF;
Alphabet make_alphabet(int n_states);

// returns false if an unknown character is met
bool emper_freq(int S, const uint8_t *chars, int64_t n, double *out);

}  // namespace ppb
