// Linear solves of the channel router's Newton iteration (tKinemat::SolveLinearSystem, src/tFlowNet/tKinemat.cpp:2085-2145:
// bandec / banbks 2154-2230 on a tridiagonal Jacobian, partial pivoting). Pure functions on plain arrays: the kernel in
// channel.cuh calls them from its threads, tests/hostcheck compiles them with g++ and loops over the lanes.
//
// Two forms:
//  * chan_band_eliminate + chan_band_back: the reference's elimination in the reference's order, one thread. About 420
//    cycles per row on B200 (a dependent FP64 operation costs ~20 cycles, the IEEE division nine of them).
//  * the partition form (chan_part_*): the rows are cut into P chunks; every chunk is eliminated by its own thread with the
//    same pivoted band elimination, for three right-hand sides (the residual and the two columns that couple the chunk to
//    its neighbours); the 2P interface unknowns (first and last level of every chunk) form a small banded system that one
//    thread solves with partial pivoting; all threads then assemble x = y - v x_left - w x_right. The chain shrinks from N
//    rows to N / P + 2P. A chunk's matrix is a principal submatrix of the Jacobian; it is pivoted like the whole one, so the
//    only requirement is that it is not singular (the Jacobian is (b > 0) on the diagonal plus an upwind difference).
//    Not the reference's order of operations: the solution differs by rounding (measured <= 1e-13 relative on the test
//    systems), which a Newton iteration that converges on |F| < 1e-4 does not see.
#pragma once

#ifndef TRIBS_HD
#ifdef __CUDACC__
#define TRIBS_HD __host__ __device__ __forceinline__
#else
#define TRIBS_HD inline
#endif
#endif

#ifdef __CUDACC__
#define CHAN_FN __host__ __device__
#else
#define CHAN_FN
#endif

// ---- the reference's order: rows 1..N in A1..A3 (row 1 already shifted left as bandec does: b, c, 0), rhs B[1..N]
CHAN_FN inline void chan_band_eliminate(double *A1, double *A2, double *A3, double *B, int N) {
  double r1 = A1[1], r2 = A2[1], r3 = A3[1], rb = B[1];
  double n1 = A1[2], n2 = A2[2], n3 = A3[2], nb = B[2];          // N >= 2
  for (int k = 1; k < N; k++) {
    const int kn = (k + 2 < N) ? k + 2 : N;                        // row after the next (its values are still the original ones)
    const double m1 = A1[kn], m2 = A2[kn], m3 = A3[kn], mb = B[kn];
    if (fabs(n1) > fabs(r1)) {                                     // partial pivoting over the two candidate rows
      double t;
      t = r1; r1 = n1; n1 = t;
      t = r2; r2 = n2; n2 = t;
      t = r3; r3 = n3; n3 = t;
      t = rb; rb = nb; nb = t;
    } else if (r1 == 0.0)
      r1 = 1.0E-20;
    const double dum = n1 / r1;
    A1[k] = r1; A2[k] = r2; A3[k] = r3; B[k] = rb;
    r1 = n2 - dum * r2;
    r2 = n3 - dum * r3;
    r3 = 0.0;
    rb = nb - dum * rb;
    n1 = m1; n2 = m2; n3 = m3; nb = mb;
  }
  if (r1 == 0.0) r1 = 1.0E-20;
  A1[N] = r1; A2[N] = 0.0; A3[N] = 0.0; B[N] = rb;
}
// x[i] = (b[i] - a2[i] x[i+1] - a3[i] x[i+2]) * R1[i], R1 = reciprocal pivots (computed by all threads beforehand)
CHAN_FN inline void chan_band_back(const double *R1, const double *A2, const double *A3, double *B, int N) {
  double x1 = B[N] * R1[N], x2 = 0.0;
  B[N] = x1;
  double c1 = R1[N - 1], c2 = A2[N - 1], c3 = A3[N - 1], cb = B[N - 1];
  for (int i = N - 1; i >= 1; i--) {
    const int ip = (i - 1 > 1) ? i - 1 : 1;
    const double p1 = R1[ip], p2 = A2[ip], p3 = A3[ip], pb = B[ip];
    double dum = cb - c2 * x1;
    if (i <= N - 2) dum -= c3 * x2;
    const double x = dum * c1;
    B[i] = x;
    x2 = x1; x1 = x;
    c1 = p1; c2 = p2; c3 = p3; cb = pb;
  }
}

// ---- partition form ---------------------------------------------------------------------------------------------
constexpr int kChanMaxParts = 32;
constexpr int kChanMinChunk = 12;      // rows per chunk below which fewer chunks are used

// Chunks against interface unknowns: P = sqrt(N / 17). Measured on 3 163 rows (profiles/r02_channel.md): 218 k cycles per
// solve at P = 13 (this rule), 241 k at P = 18, 232 k at P = 32 - a shallow optimum.
CHAN_FN inline int chan_part_count(int N) {
  int P = 1;
  while ((P + 1) * (P + 1) * 17 <= N) P++;
  const int cap = N / kChanMinChunk;
  if (P > cap) P = cap;
  return P < 1 ? 1 : (P > kChanMaxParts ? kChanMaxParts : P);
}
// chunk p covers rows first .. last (1-based, inclusive)
CHAN_FN inline void chan_part_range(int N, int P, int p, int *first, int *last) {
  const int base = N / P, extra = N % P;
  const int s = p * base + (p < extra ? p : extra);
  *first = s + 1;
  *last = s + base + (p < extra ? 1 : 0);
}

// Rows `first..last` of the tridiagonal system: sub-diagonal sub[i], diagonal dia[i], super-diagonal sup[i] (0-based i = row - 1,
// sub[0] and sup[N-1] unused), residual rhs in B (1-based like the band form). Eliminates the chunk with partial pivoting for the
// three right-hand sides and back-substitutes: on return B, V, W hold y, v, w of the chunk's rows. A1..A3 are scratch (1-based).
CHAN_FN inline void chan_part_chunk(const double *sub, const double *dia, const double *sup, double *A1, double *A2, double *A3,
                                    double *B, double *V, double *W, int first, int last, bool has_left, bool has_right) {
  const int L = last - first + 1;
  // row `first` in the shifted form (its sub-diagonal couples to the left neighbour and is the rhs V)
  double r1 = dia[first - 1], r2 = (L > 1) ? sup[first - 1] : 0.0, r3 = 0.0;
  double rb = B[first], rv = has_left ? sub[first - 1] : 0.0, rw = (L == 1 && has_right) ? sup[first - 1] : 0.0;
  for (int k = first; k < last; k++) {
    // next row k+1: (sub, dia, sup) with the super-diagonal of the chunk's last row moved to the rhs W
    const bool next_is_last = (k + 1 == last);
    double n1 = sub[k], n2 = dia[k], n3 = next_is_last ? 0.0 : sup[k];
    double nb = B[k + 1], nv = 0.0, nw = (next_is_last && has_right) ? sup[k] : 0.0;
    if (fabs(n1) > fabs(r1)) {
      double t;
      t = r1; r1 = n1; n1 = t;
      t = r2; r2 = n2; n2 = t;
      t = r3; r3 = n3; n3 = t;
      t = rb; rb = nb; nb = t;
      t = rv; rv = nv; nv = t;
      t = rw; rw = nw; nw = t;
    } else if (r1 == 0.0)
      r1 = 1.0E-20;
    const double dum = n1 / r1;
    A1[k] = 1.0 / r1; A2[k] = r2; A3[k] = r3; B[k] = rb; V[k] = rv; W[k] = rw;     // the reciprocal is off the chain
    r1 = n2 - dum * r2;
    r2 = n3 - dum * r3;
    r3 = 0.0;
    rb = nb - dum * rb;
    rv = nv - dum * rv;
    rw = nw - dum * rw;
  }
  if (r1 == 0.0) r1 = 1.0E-20;
  const double inv_last = 1.0 / r1;
  double y1 = rb * inv_last, v1 = rv * inv_last, w1 = rw * inv_last, y2 = 0.0, v2 = 0.0, w2 = 0.0;
  B[last] = y1; V[last] = v1; W[last] = w1;
  for (int i = last - 1; i >= first; i--) {
    const double inv = A1[i], c2 = A2[i], c3 = A3[i];
    double y = B[i] - c2 * y1, v = V[i] - c2 * v1, w = W[i] - c2 * w1;
    if (i <= last - 2) { y -= c3 * y2; v -= c3 * v2; w -= c3 * w2; }
    y *= inv; v *= inv; w *= inv;
    B[i] = y; V[i] = v; W[i] = w;
    y2 = y1; v2 = v1; w2 = w1;
    y1 = y; v1 = v; w1 = w;
  }
}

// The interface system: unknowns t_p = x[first_p], u_p = x[last_p], ordered t_0, u_0, t_1, u_1 ...:
//   t_p + v[first_p] u_{p-1} + w[first_p] t_{p+1} = y[first_p],   u_p + v[last_p] u_{p-1} + w[last_p] t_{p+1} = y[last_p].
// Dense Gaussian elimination with partial pivoting on at most 64 unknowns (entries outside the band stay zero); M is
// (2P) x (2P + 1) scratch, row-major with the rhs in the last column. One thread. X receives the 2P unknowns.
CHAN_FN inline void chan_part_interface(const double *B, const double *V, const double *W, int N, int P, double *M, double *X) {
  const int n = 2 * P, ld = n + 1;
  for (int i = 0; i < n * ld; i++) M[i] = 0.0;
  for (int p = 0; p < P; p++) {
    int first, last;
    chan_part_range(N, P, p, &first, &last);
    const int rt = 2 * p, ru = 2 * p + 1;
    M[rt * ld + rt] = 1.0; M[ru * ld + ru] = 1.0;
    if (first == last) {                      // a one-row chunk: t and u are the same unknown
      M[ru * ld + rt] = -1.0; M[ru * ld + n] = 0.0;
    } else {
      if (p > 0) M[ru * ld + (2 * p - 1)] = V[last];
      if (p < P - 1) M[ru * ld + (2 * p + 2)] = W[last];
      M[ru * ld + n] = B[last];
    }
    if (p > 0) M[rt * ld + (2 * p - 1)] = V[first];
    if (p < P - 1) M[rt * ld + (2 * p + 2)] = W[first];
    M[rt * ld + n] = B[first];
  }
  for (int k = 0; k < n; k++) {
    int piv = k;
    const int lim = (k + 3 < n) ? k + 3 : n;                      // the band: non-zeros of column k reach two rows below (+ fill)
    for (int i = k + 1; i < lim; i++) if (fabs(M[i * ld + k]) > fabs(M[piv * ld + k])) piv = i;
    const int cl = (k + 6 < n) ? k + 6 : n;                       // columns a row can reach after pivoting
    if (piv != k)
      for (int j = k; j <= n; j++) { if (j >= cl && j < n) continue; const double t = M[k * ld + j]; M[k * ld + j] = M[piv * ld + j]; M[piv * ld + j] = t; }
    double d = M[k * ld + k];
    if (d == 0.0) { d = 1.0E-20; M[k * ld + k] = d; }
    for (int i = k + 1; i < lim; i++) {
      const double f = M[i * ld + k] / d;
      if (f == 0.0) continue;
      for (int j = k + 1; j < cl; j++) M[i * ld + j] -= f * M[k * ld + j];
      M[i * ld + n] -= f * M[k * ld + n];
      M[i * ld + k] = 0.0;
    }
  }
  for (int k = n - 1; k >= 0; k--) {
    double s = M[k * ld + n];
    const int cl = (k + 6 < n) ? k + 6 : n;
    for (int j = k + 1; j < cl; j++) s -= M[k * ld + j] * X[j];
    X[k] = s / M[k * ld + k];
  }
}

// x of row `row` (1-based) from the chunk solution and the interface unknowns
CHAN_FN inline double chan_part_assemble(const double *B, const double *V, const double *W, const double *X, int N, int P, int row) {
  // chunk of the row
  const int base = N / P, extra = N % P;
  const int big = extra * (base + 1);
  const int r0 = row - 1;
  const int p = (r0 < big) ? r0 / (base + 1) : extra + (r0 - big) / base;
  const double left = (p > 0) ? X[2 * p - 1] : 0.0, right = (p < P - 1) ? X[2 * p + 2] : 0.0;
  return B[row] - V[row] * left - W[row] * right;
}
