"""Copy the reference's own known-answer data for Top-N scoring (tests/test_scoring.py:36-48 +
src/analytical_fm/resources/test_data/scoring/test_data.pkl: Top-1 = 0.2, Top-10 = 0.6) and for `clean_sample`
(tests/test_scoring.py:19-33) into a JSON fixture the tests can read on the GPU box.

    python tests/golden/make_scoring_fixture.py        # rewrites tests/golden/scoring.json
"""
import ast
import json
import os

import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _literal_lists(path):
    """The two string lists of test_clean_sample, read from the reference test's source (nothing is executed)."""
    tree = ast.parse(open(path).read())
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
            if name == "samples_to_clean":
                out[name] = ast.literal_eval(node.value)
            elif name == "cleaned_samples_truth":
                out[name] = ast.literal_eval(node.value.args[0])
    return out


def main():
    d = pd.read_pickle(f"{REF}/src/analytical_fm/resources/test_data/scoring/test_data.pkl")
    lit = _literal_lists(f"{REF}/tests/test_scoring.py")
    fx = {"predictions": [list(p) for p in d["predictions"]], "targets": list(d["targets"]),
          "expected": {"Top-1": 0.2, "Top-10": 0.6},
          "samples_to_clean": lit["samples_to_clean"], "cleaned_samples_truth": lit["cleaned_samples_truth"]}
    json.dump(fx, open(os.path.join(HERE, "scoring.json"), "w"), indent=0)
    print(len(fx["predictions"]), "x", len(fx["predictions"][0]), "predictions;", len(fx["samples_to_clean"]), "clean samples")


if __name__ == "__main__":
    main()
