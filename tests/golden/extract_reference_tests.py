#!/usr/bin/env python
"""Extracts the golden vectors and known answers that the reference's OWN test-suite holds for
the ASG path (``/root/reference/torch_asg/test/test_asg.py``) into
``tests/golden/reference_tests.json``.  Runs only in the build container; the JSON is committed.

Numeric list literals are lifted from the reference file with ``ast`` (no retyping); the analytic
expectations (log 2, log 32, ...) are recorded as the closed forms the reference asserts.
"""
import ast
import json
import math
import os

SRC = "/root/reference/torch_asg/test/test_asg.py"
OUT = os.path.join(os.path.dirname(__file__), "reference_tests.json")


def _lists_in(fn):
    """name -> python list for every `name = torch.tensor([...literal...])...` in a function."""
    found = {}
    for node in ast.walk(fn):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            for sub in ast.walk(node.value):
                if isinstance(sub, ast.List) and sub.elts and all(
                        isinstance(e, (ast.Constant, ast.UnaryOp)) for e in sub.elts):
                    try:
                        found.setdefault(node.targets[0].id, ast.literal_eval(sub))
                    except ValueError:
                        pass
                    break
    return found


def main():
    tree = ast.parse(open(SRC).read())
    fns = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef)}
    out = {"source": "zh217/torch-asg torch_asg/test/test_asg.py"}

    a4 = _lists_in(fns["test_asg_4"])
    out["test_asg_4"] = {
        "cite": "test_asg.py:380-464", "B": 3, "T": 5, "S": 5, "N_labels": 6,
        "inputs_BTN": a4["inputs_o"], "targets_BS": a4["targets"],
        "target_lengths": [5, 3, 4], "input_lengths": [5, 5, 5],
        "expected_loss": a4["expected_loss"],
        "expected_input_grad_BTN": a4["expected_input_grad"],
        "expected_trans_grad": a4["expected_trans_grad"],
        "tol": {"loss_abs_sum": 1e-3, "input_grad_max": 1e-4, "trans_grad_max": 1e-4},
        "note": "transition = zeros; inputs viewed (B,T,N).permute(1,0,2); reduction none, loss.sum().backward()",
    }
    f1 = _lists_in(fns["test_fac_1"])
    out["test_fac_1"] = {
        "cite": "test_asg.py:190-224", "B": 2, "T": 3, "S": 2, "N_labels": 2,
        "inputs_BTN": f1["inputs"], "targets_BS": f1["targets"],
        "expected_fac": [math.log(math.exp(1.5) + math.exp(2.5)), math.log(math.exp(2.0) + math.exp(3.0))],
        "note": "inputs are raw scores (not logs); transition zeros",
    }
    c1 = _lists_in(fns["test_fcc_1"])
    out["test_fcc_1"] = {"cite": "test_asg.py:50-72", "B": 2, "T": 3, "S": 1, "N_labels": 2,
                         "probs_BTN": c1["inputs"], "expected_fcc": [0.0, 0.0],
                         "note": "inputs = log(probs) (contains -inf); transition zeros"}
    out["test_fcc_2"] = {"cite": "test_asg.py:76-97", "B": 2, "T": 3, "N_labels": 4,
                         "note": "inputs = log(0.25) everywhere; transition zeros", "expected_fcc": [0.0, 0.0]}
    out["test_fcc_3"] = {"cite": "test_asg.py:100-124", "B": 3, "T": 300, "N_labels": 40,
                         "note": "row-normalised random probs, transition zeros => |FCC|.sum() < 1e-4"}
    out["test_fac_2"] = {"cite": "test_asg.py:228-254", "B": 1, "T": 3, "S": 2, "N_labels": 4,
                         "targets_BS": [0, 1], "expected_fac": [-math.log(32.0)],
                         "note": "inputs = log(0.25) everywhere; transition zeros"}
    a1 = _lists_in(fns["test_asg_1"])
    out["test_asg_1"] = {"cite": "test_asg.py:292-321", "B": 2, "T": 3, "S": 2, "N_labels": 2,
                         "probs_BTN": a1["inputs"], "targets_BS": [0, 1, 0, 1],
                         "expected_loss": [math.log(2.0), 0.0], "note": "inputs = log(probs); reduction none"}
    out["test_asg_2"] = {"cite": "test_asg.py:325-351", "B": 1, "T": 3, "S": 2, "N_labels": 4,
                         "targets_BS": [0, 1], "expected_loss_mean": math.log(32.0),
                         "note": "inputs = log(0.25); reduction mean"}
    out["test_asg_3"] = {"cite": "test_asg.py:355-376", "B": 1, "T": 3, "S": 4, "N_labels": 4,
                         "targets_BS": [0, 1, 1, 1],
                         "note": "target longer than T: truncated to first T labels; two calls must agree"}
    out["test_fac_grad"] = {"cite": "test_asg.py:258-288", "B": 3, "T": 5, "S": 3, "N_labels": 3,
                            "targets_BS": [1, 2, 1, 0, 1, 0, 1, 0, 0], "target_lengths": [3, 2, 1],
                            "note": "gradcheck with variable target lengths"}
    json.dump(out, open(OUT, "w"), indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
