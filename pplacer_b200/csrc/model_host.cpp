// See model_host.hpp.  Host C++ only.
#include "model_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace ppb {

// ---- incomplete gamma ---------------------------------------------------------------------
// P(a,x) = gamma(a,x)/Gamma(a).  Power series for x < a+1, modified-Lentz continued fraction for
// the complement otherwise; evaluated in long double so the double result is correctly rounded
// to within an ulp or two, which is what the Yang-1994 rates need (gamma.ml:14-40 call sites).
static long double gamma_p_ld(long double a, long double x) {
  if (!(x > 0.0L)) return 0.0L;
  const long double lead = expl(a * logl(x) - x - lgammal(a));
  if (x < a + 1.0L) {
    long double term = 1.0L / a, sum = term, ap = a;
    for (int n = 1; n < 200000; ++n) {
      ap += 1.0L;
      term *= x / ap;
      sum += term;
      if (fabsl(term) <= fabsl(sum) * 1e-22L) break;
    }
    return lead * sum;
  }
  const long double fpmin = std::numeric_limits<long double>::min() / std::numeric_limits<long double>::epsilon();
  long double b = x + 1.0L - a, c = 1.0L / fpmin, d = 1.0L / b, h = d;
  for (int i = 1; i < 200000; ++i) {
    const long double an = -static_cast<long double>(i) * (static_cast<long double>(i) - a);
    b += 2.0L;
    d = an * d + b;
    if (fabsl(d) < fpmin) d = fpmin;
    c = b + an / c;
    if (fabsl(c) < fpmin) c = fpmin;
    d = 1.0L / d;
    const long double delta = d * c;
    h *= delta;
    if (fabsl(delta - 1.0L) <= 1e-22L) break;
  }
  return 1.0L - lead * h;
}

double gamma_p(double a, double x) { return static_cast<double>(gamma_p_ld(a, x)); }

double gamma_p_inv(double p, double a) {
  if (!(p > 0.0)) return 0.0;
  if (p >= 1.0) return std::numeric_limits<double>::infinity();
  // bracket, then Newton steps kept inside the bracket (falls back to bisection)
  long double lo = 0.0L, hi = std::max(1.0L, static_cast<long double>(a));
  while (gamma_p_ld(a, hi) < p) { lo = hi; hi *= 2.0L; }
  long double x = 0.5L * (lo + hi);
  const long double lg = lgammal(static_cast<long double>(a));
  for (int it = 0; it < 400; ++it) {
    const long double f = gamma_p_ld(a, x) - static_cast<long double>(p);
    (f > 0.0L ? hi : lo) = x;
    const long double dens = expl((static_cast<long double>(a) - 1.0L) * logl(x) - x - lg);
    long double nx = dens > 0.0L ? x - f / dens : -1.0L;
    if (!(nx > lo && nx < hi)) nx = 0.5L * (lo + hi);
    const bool done = fabsl(nx - x) <= 1e-19L * fabsl(x);
    x = nx;
    if (done) break;
  }
  return static_cast<double>(x);
}

// Gamma.discrete_gamma (gamma.ml:52-70): category boundaries at the i/n quantiles (eq. 9 of Yang
// 1994), category means via P(alpha+1, cut) differences (eq. 10), rescaled to mean one
// (make_mean_one, gamma.ml:27-30).  inverse_gamma_cdf's small-y guard (gamma.ml:14-25) is kept.
void discrete_gamma(int n_cat, double alpha, double *out) {
  std::vector<double> cum(static_cast<size_t>(n_cat) + 1);
  cum[0] = 0.0;
  cum[n_cat] = 1.0;
  for (int i = 1; i < n_cat; ++i) {
    const double y = static_cast<double>(i) / static_cast<double>(n_cat);
    const double cut = (y < gamma_p(alpha, 1e-7)) ? 0.0 : gamma_p_inv(y, alpha);
    cum[i] = gamma_p(alpha + 1.0, cut);
  }
  double tot = 0.0;
  for (int i = 0; i < n_cat; ++i) { out[i] = cum[i + 1] - cum[i]; tot += out[i]; }
  const double n = static_cast<double>(n_cat);
  for (int i = 0; i < n_cat; ++i) out[i] = n * out[i] / tot;
}

// ---- symmetric eigenproblem ---------------------------------------------------------------
// Cyclic Jacobi with the classical stable rotation formulas; plenty for S <= 20.
void jacobi_eigen(int n, const double *sym, double *eval, double *evec) {
  std::vector<double> a(sym, sym + static_cast<size_t>(n) * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) evec[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) off += a[p * n + q] * a[p * n + q];
    if (off == 0.0 || off < 1e-300) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = a[p * n + q];
        if (apq == 0.0) continue;
        const double tau = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < n; ++k) {  // A <- A J
          const double kp = a[k * n + p], kq = a[k * n + q];
          a[k * n + p] = c * kp - s * kq;
          a[k * n + q] = s * kp + c * kq;
        }
        for (int k = 0; k < n; ++k) {  // A <- J^T A
          const double pk = a[p * n + k], qk = a[q * n + k];
          a[p * n + k] = c * pk - s * qk;
          a[q * n + k] = s * pk + c * qk;
        }
        for (int k = 0; k < n; ++k) {  // V <- V J
          const double kp = evec[k * n + p], kq = evec[k * n + q];
          evec[k * n + p] = c * kp - s * kq;
          evec[k * n + q] = s * kp + c * kq;
        }
      }
  }
  for (int i = 0; i < n; ++i) eval[i] = a[i * n + i];
}

// Diagd (diagd.ml): B gets diagonal -(1/pi_i) sum_{k != i} r_ki pi_k (:140-152); with D = diag(pi),
// U Lambda U^T = D^1/2 B D^1/2, x = D^-1/2 U, xit = D^1/2 U (:123-132); finally lambda is scaled
// by 1/rate with rate = -sum_i Q_ii pi_i (:158-167), Q = x diag(lambda) xit^T (linear_c.c:92).
void diagd_normed(int S, const double *exch, const double *pi, double *x, double *lambda, double *xit) {
  const size_t nn = static_cast<size_t>(S) * S;
  std::vector<double> b(nn), m(nn), u(nn);
  auto inv0 = [](double v) { return v == 0.0 ? 1.0 : 1.0 / v; };  // safe_invert diagd.ml:118
  for (int i = 0; i < S; ++i)
    for (int j = 0; j < S; ++j) {
      if (i != j) { b[i * S + j] = exch[i * S + j]; continue; }
      double tot = 0.0;
      for (int k = 0; k < S; ++k)
        if (k != i) tot += exch[k * S + i] * pi[k];
      b[i * S + j] = -(tot * inv0(pi[i]));
    }
  std::vector<double> root(S);
  for (int i = 0; i < S; ++i) root[i] = std::sqrt(pi[i]);
  for (int i = 0; i < S; ++i)
    for (int j = 0; j < S; ++j) m[i * S + j] = root[i] * (b[i * S + j] * root[j]);
  for (int i = 0; i < S; ++i)
    for (int j = i + 1; j < S; ++j) m[i * S + j] = m[j * S + i] = 0.5 * (m[i * S + j] + m[j * S + i]);
  jacobi_eigen(S, m.data(), lambda, u.data());
  for (int i = 0; i < S; ++i)
    for (int j = 0; j < S; ++j) {
      x[i * S + j] = inv0(root[i]) * u[i * S + j];
      xit[i * S + j] = root[i] * u[i * S + j];
    }
  double rate = 0.0;
  for (int i = 0; i < S; ++i) {
    double qii = 0.0;
    for (int k = 0; k < S; ++k) qii += lambda[k] * x[i * S + k] * xit[i * S + k];
    rate -= qii * pi[i];
  }
  const double sc = 1.0 / rate;
  for (int i = 0; i < S; ++i) lambda[i] *= sc;
}

// ---- alphabets ----------------------------------------------------------------------------
static const char kProt[] = "ARNDCQEGHILKMFPSTWYV";  // prot_models.ml:5

Alphabet make_alphabet(int S) {
  Alphabet al;
  al.n_states = S;
  std::memset(al.code_of_char, kBadCode, sizeof al.code_of_char);
  std::memset(al.mask_of_code, 0, sizeof al.mask_of_code);
  if (S == 4) {
    // code = 4-bit state mask itself (A=1, C=2, G=4, T=8); nuc_models.ml:28-45
    al.n_codes = 16;
    for (int c = 0; c < 16; ++c) al.mask_of_code[c] = static_cast<uint32_t>(c);
    const struct { char ch; const char *set; } tab[] = {
        {'A', "A"}, {'C', "C"}, {'G', "G"}, {'T', "T"}, {'R', "AG"}, {'Y', "CT"}, {'M', "AC"}, {'K', "GT"},
        {'S', "CG"}, {'W', "AT"}, {'H', "ACT"}, {'B', "CGT"}, {'V', "ACG"}, {'D', "AGT"}, {'N', "ACGT"}, {'X', "ACGT"}};
    for (const auto &t : tab) {
      uint8_t m = 0;
      for (const char *p = t.set; *p; ++p) m |= static_cast<uint8_t>(1u << (std::strchr("ACGT", *p) - "ACGT"));
      al.code_of_char[static_cast<unsigned char>(t.ch)] = m;
    }
    al.any_code = 15;
  } else {
    // codes 0..19 one-hot, 20 = B (D|N), 21 = J (I|L), 22 = Z (E|Q), 23 = any; prot_models.ml:7-44
    al.n_codes = 24;
    for (int i = 0; i < 20; ++i) {
      al.mask_of_code[i] = 1u << i;
      al.code_of_char[static_cast<unsigned char>(kProt[i])] = static_cast<uint8_t>(i);
    }
    auto bit = [](char c) { return 1u << (std::strchr(kProt, c) - kProt); };
    al.mask_of_code[20] = bit('D') | bit('N');
    al.mask_of_code[21] = bit('I') | bit('L');
    al.mask_of_code[22] = bit('E') | bit('Q');
    al.mask_of_code[23] = (1u << 20) - 1u;
    al.code_of_char[static_cast<unsigned char>('B')] = 20;
    al.code_of_char[static_cast<unsigned char>('J')] = 21;
    al.code_of_char[static_cast<unsigned char>('Z')] = 22;
    al.code_of_char[static_cast<unsigned char>('X')] = 23;
    al.any_code = 23;
  }
  // '-' and '?' carry the all-ones vector but are uninformative (alignment.ml:235)
  al.code_of_char[static_cast<unsigned char>('-')] = static_cast<uint8_t>(al.any_code | kGapBit);
  al.code_of_char[static_cast<unsigned char>('?')] = static_cast<uint8_t>(al.any_code | kGapBit);
  return al;
}

// Alignment.emper_freq (alignment.ml:186-217): every character except '-' and '?' adds its
// l1-normalised indicator vector ('X' included); the total is l1-normalised.
bool emper_freq(int S, const uint8_t *chars, int64_t n, double *out) {
  const Alphabet al = make_alphabet(S);
  std::vector<int64_t> count(kMaxCodes, 0);
  for (int64_t i = 0; i < n; ++i) {
    const uint8_t c = chars[i];
    if (c == '-' || c == '?') continue;
    const uint8_t code = al.code_of_char[c];
    if (code == kBadCode) return false;
    ++count[code & 0x7f];
  }
  // accumulate in character order is not reproducible from counts, but each normalised vector is
  // a multiple of 1/popcount, so summing counts per code first changes nothing beyond rounding of
  // the (exactly representable for popcount 1,2,4) increments; popcount-3 codes are summed last.
  for (int s = 0; s < S; ++s) out[s] = 0.0;
  for (int code = 0; code < al.n_codes; ++code) {
    if (!count[code]) continue;
    const uint32_t m = al.mask_of_code[code];
    const int pc = __builtin_popcount(m);
    const double w = static_cast<double>(count[code]) * (1.0 / static_cast<double>(pc));
    for (int s = 0; s < S; ++s)
      if (m >> s & 1u) out[s] += w;
  }
  double tot = 0.0;
  for (int s = 0; s < S; ++s) tot += out[s];
  const double sc = 1.0 / tot;
  for (int s = 0; s < S; ++s) out[s] *= sc;
  return true;
}

}  // namespace ppb
