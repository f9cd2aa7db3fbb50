"""rehuel_b200 -- B200-native batched ensemble integrator behind Rehuel's irk API.

The product is ``librehuel_b200.so`` (rehuel_b200/csrc: CUDA for sm_100a + C++ host layer, C ABI
in include/rehuel_b200.h).  This package only binds it (``capi``), wraps device-resident calls for
torch tensors (``device``) and generates the synthetic benchmark ensembles (``ensembles``).
Importing it without the built library raises ImportError: there is no CPU fallback.
"""
from . import capi, ensembles  # noqa: F401
from .capi import (Options, RehuelError, default_options, irk_ensemble, irk_single, make_options,  # noqa: F401
                   method_to_name, name_to_method, name_to_problem)

__all__ = ["capi", "ensembles", "Options", "RehuelError", "default_options", "irk_ensemble", "irk_single",
           "make_options", "method_to_name", "name_to_method", "name_to_problem"]
