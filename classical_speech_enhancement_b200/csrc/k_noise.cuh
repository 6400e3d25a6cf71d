#pragma once
#include "cse_common.cuh"
