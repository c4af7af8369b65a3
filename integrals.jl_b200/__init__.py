"""integrals_b200 -- B200-native batched quadrature engine behind the Integrals.jl solve interface.

Layout: csrc/ (CUDA kernels + the C ABI of include/integrals_b200.h -> libintegrals_b200.so),
_lib.py (ctypes binding), integrands.py (registered device-side integrand markers), interface.py
(host-side mirror of the reference's problem / algorithm / solve interface) and sharding.py
(one process per GPU: batches of independent problems split across ranks).  There is no CPU
fallback: without the built library and a B200 the solve calls raise.
"""
from ._lib import (Engine, IB200Error, FAMILY_IDS, TRANSFORM_IDS, RULE_IDS, SYMBOLS, LIB_PATH,
                   MAXEVALS_UNBOUNDED, default_engine, load_library)
from .build import build_library
from .integrands import (RegisteredIntegrand, SinXP, GenzOscillatory, GenzProductPeak, GenzCornerPeak, GenzGaussian,
                         GenzC0, GenzDiscontinuous, CosProduct, GaussPoly, ExpLinear, GLTest, VectorOf, ALL_INTEGRANDS)
from .interface import (ParameterSensitivities, DomainError, retcode_of, IntegralProblem, IntegralFunction, BatchIntegralFunction, CudaIntegralFunction, SampledIntegralProblem, IntegralSolution, IntegralCache, SampledIntegralCache,
                        Range, ReturnCode, CommonKwargError, ChangeOfVariables, transformation_if_inf,
                        transformation_tan_inf, transformation_cot_inf, QuadGKB200, HCubatureB200, QuadratureRuleB200,
                        GaussLegendreB200, TrapezoidalB200, SimpsonsB200, init, solve, solve_cache)
from .sharding import shard_range, shard_indices, solve_sharded, solve_sampled_sharded, solve_sampled_split_axis
