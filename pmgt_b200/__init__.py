"""pmgt_b200 -- B200-native (sm_100a) implementation of PMGT's pre-training hot path.

Drop-in names of the reference (uoo723/PMGT):

    pmgt.pmgt.configuration_pmgt.PMGTConfig  -> pmgt_b200.PMGTConfig
    pmgt.pmgt.datasets.{PMGTDataset, pmgt_collate_fn, get_input_tensor}
    pmgt.pmgt.modeling_pmgt.{PMGTModel, PMGTGraphConstructLoss, PMGTNodeConstructLoss, ...}
    pmgt.pmgt.models.PMGT
    pmgt.pmgt.trainer  (module)              -> pmgt_b200.trainer
    pmgt.optimizers.DenseSparseAdamW

All compute goes through ``libpmgt_b200.so`` (hand-written CUDA, C ABI in
``include/pmgt_b200.h``); there is no CPU fallback.
"""
from .configuration_pmgt import PMGTConfig  # noqa: F401
from .graph import ItemGraph  # noqa: F401


def __getattr__(name):
    # torch-dependent modules are imported lazily so that `import pmgt_b200` stays cheap
    import importlib

    table = {
        "PMGT": ".models", "PMGTModel": ".modeling_pmgt", "PMGTForPreTrainingOutput": ".modeling_pmgt",
        "PMGTGraphConstructLoss": ".modeling_pmgt", "PMGTNodeConstructLoss": ".modeling_pmgt",
        "PMGTDataset": ".datasets", "pmgt_collate_fn": ".datasets", "get_input_tensor": ".datasets",
        "DenseSparseAdamW": ".optimizers",
    }
    if name in table:
        return getattr(importlib.import_module(table[name], __name__), name)
    if name in ("trainer", "datasets", "models", "modeling_pmgt", "optimizers", "ops", "synthetic", "utils", "build"):
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
