// Frozen-phi test chains (LabeledLDA.py:179-212).  PLACEHOLDER.
#pragma once
#include <stdint.h>
static int test_chains_run(int, int, double, const double *, long long, const int64_t *, const int32_t *, const int32_t *, const int32_t *, int, int, uint64_t, double *) { return 1; }
