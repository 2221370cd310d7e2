"""Regenerates tests/golden/reference_suite_anchors.json from the CPU oracle (run from the repo root).
The reference (Scala/Spark) cannot run in this image, so these are oracle-emitted regression anchors."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O  # noqa: E402

x1, y = O.generate_gd_input(2.0, -1.5, 10000, 42)
X = np.stack([np.ones_like(x1), x1], axis=1)
D = O.Data(y, X=X)
r = O.agd_run(D, "logistic", "simple", [1.0, -1.0], convergence_tol=1e-12, num_iterations=10)
_, lg = O.gd_run(D, "logistic", "simple", [1.0, -1.0], step_size=1.0, num_iterations=50)
r2 = O.agd_run(D, "logistic", "squared_l2", [0.3, 0.12], convergence_tol=1e-12, num_iterations=10, reg_param=0.2)
G = {"_note": "produced by oracle/agd_oracle.c (tests/golden/make_anchors.py); regression anchors, NOT reference-emitted goldens",
     "T1_agd_loss_history": r.loss_history.tolist(), "T1_agd_weights": r.weights.tolist(),
     "T1_gd_last_loss": float(lg[-1]), "T2_agd_weights": r2.weights.tolist(),
     "T2_agd_last_loss": float(r2.loss_history[-1])}
json.dump(G, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_suite_anchors.json"), "w"), indent=1)
