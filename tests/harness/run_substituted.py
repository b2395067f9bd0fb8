"""Run the reference's partial_modality_training.py with its inline model class rebound to the B200 head.

    python run_substituted.py <path to partial_modality_training.py>

The scripts define their models inline, so the head cannot be swapped by import (SURVEY.md 8b "Model boundary").  This
runner parses the UNCHANGED script, replaces the ``class PartialModalityNet`` statement by an import of
``multimodal_survival_prediction_b200.head.PartialModalityNet`` and ``def gate_entropy_loss`` by the fused one, and executes
everything else as written: datasets, loaders, train/validate loops, clipping, Adam, scheduler, checkpoints, JSON summary.
"runs unchanged" is claimed for the loss/metric (tests/test_ref_scripts_gpu.py runs the file itself); for the model it
is by this substitution."""
import ast
import sys


def main(path):
    with open(path, encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=path)
    swapped = []
    for i, node in enumerate(tree.body):
        if isinstance(node, ast.ClassDef) and node.name == "PartialModalityNet":
            tree.body[i] = ast.parse("from multimodal_survival_prediction_b200.head import PartialModalityNet").body[0]
            swapped.append("PartialModalityNet")
        elif isinstance(node, ast.FunctionDef) and node.name == "gate_entropy_loss":
            tree.body[i] = ast.parse("from multimodal_survival_prediction_b200.head import gate_entropy_loss").body[0]
            swapped.append("gate_entropy_loss")
    assert swapped == ["PartialModalityNet", "gate_entropy_loss"], swapped
    ast.fix_missing_locations(tree)
    print("harness: rebound", ", ".join(swapped), "to multimodal_survival_prediction_b200.head", flush=True)
    exec(compile(tree, path, "exec"), {"__name__": "__main__", "__file__": path})


if __name__ == "__main__":
    main(sys.argv[1])
