"""Pull the pure-torch definitions of the TPDM hot path straight out of /root/reference by AST (no diffusers,
no pyrootutils needed) so the oracle restatements can be pinned against the reference's own code.

Only usable in the build container (where /root/reference exists).  Nothing under tests -m gpu, smoke() or
bench.py calls this; `oracle/make_golden.py` does, and commits the resulting vectors to tests/golden/.
"""
from __future__ import annotations

import ast
import math
import os
from typing import Optional

REF = os.environ.get("TPDM_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "src/models/stable_diffusion_3/modeling_sd3_pnt.py"))


def _exec_nodes(path: str, names, env: dict) -> dict:
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, env)
    missing = [n for n in names if n not in env]
    if missing:
        raise RuntimeError(f"reference definitions not found in {path}: {missing}")
    return env


def load_reference_pieces() -> dict:
    """Returns {'reshape_hidden_states_to_2d', 'CustomAdaGroupNormZeroSingle', 'TimePredictor',
    'custom_step_body', 'get_ref_beta', 'get_kl_beta'} executed from the reference's source text."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    env = {"torch": torch, "nn": nn, "F": F, "Optional": Optional, "math": math}
    _exec_nodes(os.path.join(REF, "src/models/stable_diffusion_3/modeling_sd3_pnt.py"),
                ["reshape_hidden_states_to_2d", "CustomAdaGroupNormZeroSingle", "TimePredictor"], env)

    # custom_step is a method of a diffusers subclass; lift its body (model_utilis.py:52-74) into a free function.
    src = open(os.path.join(REF, "src/models/model_utilis.py")).read()
    tree = ast.parse(src)
    fn = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "custom_step":
            fn = node
    if fn is None:
        raise RuntimeError("custom_step not found")
    fn.returns = None
    for a in fn.args.args:
        a.annotation = None
    fn.name = "custom_step_body"
    env2 = {"torch": torch, "CustomFlowMatchEulerDiscreteSchedulerOutput": lambda prev_sample: (prev_sample,)}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "model_utilis.py", "exec"), env2)
    env["custom_step_body"] = lambda mo, sn, s, x: env2["custom_step_body"](None, mo, sn, s, x, return_dict=False)[0]

    env3 = {"torch": torch, "math": math, "EPSILON": 1e-3, "CONCENTRATION": 20, "ex": math.exp(1)}
    _exec_nodes(os.path.join(REF, "src/models/reference_distributions.py"), ["get_ref_beta"], env3)
    env["get_ref_beta"] = env3["get_ref_beta"]
    env4 = {"torch": torch}
    _exec_nodes(os.path.join(REF, "src/train/train_utilis.py"), ["get_kl_beta"], env4)
    env["get_kl_beta"] = env4["get_kl_beta"]
    return env
