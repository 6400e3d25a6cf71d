#pragma once
#include "cse_common.cuh"
static void cse_fill_resampler(CseTables* t) { (void)t; }
