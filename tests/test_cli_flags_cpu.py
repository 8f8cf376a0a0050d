"""The root `reconstruct.py` accepts every flag of the reference's CLI with the same default (reconstruct.py:7-141 of the
reference): its own `parse_args` is executed here (only that function: the module imports the trainers) and compared."""
import ast
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference/reconstruct.py")


def test_every_reference_flag_with_its_default(monkeypatch):
    if not REF.exists():
        pytest.skip("reference tree not present (GPU box)")
    tree = ast.parse(REF.read_text())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "parse_args")
    mod = ast.Module(body=[ast.Import(names=[ast.alias(name="argparse")]), ast.Import(names=[ast.alias(name="ast")]), fn],
                     type_ignores=[])
    ns = {}
    exec(compile(ast.fix_missing_locations(mod), str(REF), "exec"), ns)
    monkeypatch.setattr(sys, "argv", ["reconstruct.py"])
    ref = vars(ns["parse_args"]())
    sys.path.insert(0, str(ROOT))
    import reconstruct as cli

    ours = vars(cli.parse_args([]))
    assert set(ref) <= set(ours), set(ref) - set(ours)
    assert set(ours) - set(ref) == {"plms_state", "honour_num_inference_steps", "shard"}
    assert {k: ours[k] for k in ref} == ref
    # the literal-typed flags parse like the reference's
    a = cli.parse_args(["--image_roi", "[160,160,128]", "--latent_pad", "(1,1,1,1,0,0)", "--beta_start", "0.0015"])
    assert a.image_roi == [160, 160, 128] and a.latent_pad == (1, 1, 1, 1, 0, 0) and a.beta_start == 0.0015
