"""Prints bench.py's geometry_rates() block (SURVEY f2 slice) as one JSON line - the same numbers `bench.py` puts under
config.extra.geometry_tail_N32, without running the whole bench."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

if __name__ == "__main__":
    hbm = bench.load_peaks()[1]
    print(json.dumps(bench.geometry_rates(torch.device("cuda:0"), hbm)))
