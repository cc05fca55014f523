"""sharp_b200 -- B200-native (sm_100a) implementation of SHARP's ensemble random-projection clustering path.

Host side: a Python mirror of the reference's R operator API (SHARP(), SHARP_unlimited*, RPmat, ranM,
get_opt_hclust, getrowColor, wMetaC, sMetaC ...) over the C ABI of ``libsharpb200.so`` (include/sharp_b200.h).
R is not available in this image; INTEGRATION.md shows the `.Call` glue a maintainer of the reference adds.
Importing the package does not need a GPU; every compute call does (there is no CPU fallback).
"""
import os as _os

# sharp_run_parts drives ~10 CUDA streams; with the default 8 hardware queues streams alias and copies / kernels of
# different parts serialise on false dependencies.  Read by the CUDA driver at initialisation: set before first use.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import _lib
from ._lib import Context, HcParams, RStop, RunParams, SharpError, device_count, device_info, hc_params
from .rrng import r_sample_perm
from .api import (ARI, SHARP, Expression, RPmat, SHARP_fpart, SHARP_large, SHARP_small, SHARP_unlimited,
                  SHARP_unlimited2, SHARP_unlimited3, colorL, get_context, get_opt_hclust, getA, getrowColor, getss,
                  ranM, ranM2, run_Mtimes_SHARP, sMetaC, set_devices, set_verbose, testlog, wMetaC)

__all__ = ["_lib", "Context", "HcParams", "RunParams", "SharpError", "RStop", "device_count", "device_info",
           "hc_params", "ranM", "ranM2", "r_sample_perm", "SHARP", "SHARP_small", "SHARP_large", "SHARP_unlimited",
           "SHARP_unlimited2", "SHARP_fpart", "SHARP_unlimited3", "RPmat", "get_opt_hclust", "getrowColor", "wMetaC",
           "sMetaC", "getA", "getss", "testlog", "ARI", "run_Mtimes_SHARP", "Expression", "get_context", "set_devices",
           "set_verbose", "colorL"]
