#!/bin/bash
# marginals on a graph whose factorisation ends in a tail chain (garage fixture, d = 6): the chain's inverse diagonal blocks
# reach the sparse inverse through the optional second copy
cd tests && timeout 100 python - <<'PY'
import sys
sys.path.insert(0, "..")
import test_gpu_parity as t
import numpy as np
opt, fx = t._product_from_fixture("garage")
opt._ensure_uploaded(); ctx = opt.context; assert ctx.build_structure()
print("chain links:", ctx.factor_info()["chain_links"])
t.test_marginals_match_dense_inverse_of_the_oracle_hessian.__wrapped__("garage") if hasattr(t.test_marginals_match_dense_inverse_of_the_oracle_hessian, "__wrapped__") else t.test_marginals_match_dense_inverse_of_the_oracle_hessian("garage")
print("garage marginals ok")
PY
