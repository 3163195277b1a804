"""partapp_b200 -- B200-native pictorial-structures inference (the `--find_obj` hot path of partapp).

The package is a thin host layer over libpsinfer.so (hand-written sm_100a CUDA kernels behind the C ABI of
include/psinfer.h).  There is no CPU implementation in here: without the built extension and a GPU nothing computes.
"""
from . import capi
from .capi import PsInferError, load_library
from .objectdetect import (ExpParam, Joint, PartConf, PsContext, RootPosteriorResult, computeRootPosteriorRot,
                           computeRotJointMarginal, findLocalMax, getMaxStates, rot_from_index, scale_from_index)

__all__ = ["capi", "PsInferError", "load_library", "ExpParam", "Joint", "PartConf", "PsContext",
           "RootPosteriorResult", "computeRootPosteriorRot", "computeRotJointMarginal", "findLocalMax",
           "getMaxStates", "rot_from_index", "scale_from_index"]
