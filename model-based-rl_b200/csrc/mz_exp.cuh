// exp() used for Node.expand priors (mcts.py:52: math.exp(logit) -> libm exp on the host).
#pragma once
#include "mz_common.cuh"

MZ_DEV double mz_exp(double x) { return exp(x); }
