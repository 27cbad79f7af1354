// ref_regression.cpp -- TEST INFRASTRUCTURE (CPU baseline of BASELINE.json's second metric).
//
// Drives the reference's OWN Regression.h / Matrix.cpp / FHE-SI code (compiled by oracle/build_ref.py
// against the NTL stand-in) through the phases Test_Regression.cpp:24-63 times, with two things the
// reference's driver cannot do: the parameters (logQ, xi) are given explicitly -- the reference sizes
// them from the one file it is handed (Test_Regression.cpp:100-108), but a sharded run must size them
// from the GLOBAL N -- and only the first `maxBlocks` blocks of the file are
// processed, so that the N-proportional phases (batch, encryption, data-phase products) can be sampled
// at full-size parameters in bounded time and scaled by the block count, while the N-independent
// phases (set-up = key generation, the key-switch / rotation / adjugate tail, decryption) run whole.
// Our own driver code, written against the reference's public API; wall clock, not clock().
//
//   ref_regression <datafile> <p> <g> <logQ> <xi> <maxBlocks|0> [seed] [skipTail]  -> one JSON line
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "Ciphertext.h"
#include "FHE-SI.h"
#include "FHEContext.h"
#include "Regression.h"

static double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
  if (argc < 7) {
    fprintf(stderr, "usage: ref_regression datafile p g logQ xi maxBlocks [seed] [skipTail]\n");
    return 2;
  }
  const std::string datafile = argv[1];
  const unsigned p = atoi(argv[2]), g = atoi(argv[3]), logQ = atoi(argv[4]);
  const long xi = atol(argv[5]);
  const unsigned maxBlocks = atoi(argv[6]);
  const long seed = argc > 7 ? atol(argv[7]) : 12345;
  const bool skipTail = argc > 8 && atoi(argv[8]) != 0;
  srand48(seed);
  SetSeed(to_ZZ(seed));

  unsigned blockSize = 1;
  for (unsigned val = (p - 1) / 2 - 1; val > 1; val >>= 1) blockSize <<= 1;  // Test_Regression.cpp:86-91
  Matrix<ZZ> all, rawData;
  vector<ZZ> allLabels, labels;
  unsigned dim = 0;
  if (!LoadData(all, allLabels, dim, datafile)) return 1;
  const unsigned fileRows = all.NumRows();
  const unsigned fileBlocks = (fileRows + blockSize - 1) / blockSize;
  unsigned rows = fileRows;
  if (maxBlocks && maxBlocks * blockSize < rows) rows = maxBlocks * blockSize;
  for (unsigned i = 0; i < rows; ++i) {
    vector<ZZ> r(dim);
    for (unsigned j = 0; j < dim; ++j) r[j] = all[i][j];
    rawData.AddRow(r);
    labels.push_back(allLabels[i]);
  }
  const unsigned nBlocks = (rows + blockSize - 1) / blockSize;

  FHEcontext context(p - 1, logQ, p, g, 3);
  activeContext = &context;
  context.SetUpSIContext(xi);

  vector<ZZ> thetaPT;
  ZZ detPT;
  RegressPT(thetaPT, detPT, rawData, labels);

  const double t0 = now();
  Regression regress(context);  // :25  keys, s^2 and rotation key-switch matrices
  const double tSetup = now();
  vector<vector<Plaintext>> ptxtData;
  vector<Plaintext> ptxtLabels;
  BatchData(ptxtData, ptxtLabels, rawData, labels, context);  // :31
  const double tBatch = now();
  regress.AddData(ptxtData, ptxtLabels);  // :36
  const double tEnc = now();
  // The N-proportional part of Regress(), on its own: the products X^T y and X X^T summed in tensor form
  // (Regression.h:103-107, Matrix.cpp:80-97,149-173), on a second encryption of the same blocks (the
  // class keeps its data private).  regression_s - data_phase_s is then the N-independent tail.
  double dataPhase = 0;
  {
    FHESIPubKey &pk = regress.GetPublicKey();
    Matrix<Ciphertext> X((Ciphertext(pk)));
    vector<Ciphertext> y;
    for (unsigned i = 0; i < ptxtData.size(); ++i) {
      vector<Ciphertext> row(ptxtData[i].size(), Ciphertext(pk));
      for (unsigned j = 0; j < ptxtData[i].size(); ++j) pk.Encrypt(row[j], ptxtData[i][j]);
      Ciphertext l(pk);
      pk.Encrypt(l, ptxtLabels[i]);
      X.AddRow(row);
      y.push_back(l);
    }
    const double d0 = now();
    Matrix<Ciphertext> Xc = X;
    Xc.Transpose();
    Matrix<Ciphertext> last = Xc * y;
    Xc.MultByTranspose();
    dataPhase = now() - d0;
  }
  const double tEnc2 = now();
  vector<Ciphertext> encTheta;
  Ciphertext encDet(regress.GetPublicKey());
  double tReg = tEnc2, tDec = tEnc2;
  bool ok = true;
  std::string got = "[";
  if (!skipTail) {
    regress.Regress(encTheta, encDet);  // :43
    tReg = now();
    tDec = tReg;
    FHESISecKey secretKey = regress.GetSecretKey();
    Plaintext tmp(context);
    vector<ZZ_pX> msgs;
    for (unsigned i = 0; i <= encTheta.size(); ++i) {  // :50-60
      secretKey.Decrypt(tmp, i < encTheta.size() ? encTheta[i] : encDet);
      tmp.DecodeSlots(msgs);
      const ZZ want = (i < encTheta.size() ? thetaPT[i] : detPT) % to_ZZ(p);
      const long v = deg(msgs[0]) < 0 ? 0 : to_long(rep(coeff(msgs[0], 0)));
      ok = ok && to_ZZ(v) == want;
      got += (i ? ", " : "") + std::to_string(v);
    }
    tDec = now();
  }
  got += "]";
  printf("{\"file_rows\": %u, \"file_blocks\": %u, \"rows\": %u, \"blocks\": %u, \"dim\": %u, \"p\": %u, \"g\": %u, "
         "\"logQ\": %u, \"xi\": %ld, \"skip_tail\": %s, \"setup_s\": %.6f, \"batch_s\": %.6f, \"encryption_s\": %.6f, "
         "\"data_phase_s\": %.6f, \"regression_s\": %.6f, \"decryption_s\": %.6f, \"total_s\": %.6f, \"decrypted\": %s, "
         "\"correct\": %s}\n",
         fileRows, fileBlocks, rows, nBlocks, dim, p, g, logQ, xi, skipTail ? "true" : "false", tSetup - t0,
         tBatch - tSetup, tEnc - tBatch, dataPhase, tReg - tEnc2, tDec - tReg, (tEnc - t0) + (tDec - tEnc2), got.c_str(),
         (ok || skipTail) ? "true" : "false");
  return ok ? 0 : 1;
}
