"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference algorithm).

PARITY UNPINNED: the reference (yanwu2014/swne) has no tests or golden vectors for the NMF path and
the arithmetic lives in the un-vendored dependency linxihui/NNLM (unpinned GitHub HEAD, 0.4.x line;
/root/reference/DESCRIPTION:42-43).  Neither R nor NNLM is runnable in the build container, so this
oracle restates NNLM's published algorithm and the in-tree R glue
(/root/reference/R/nmf.R, R/swne_plotting.R) and is pinned only by independent cross-checks.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  swne_b200/ never does.
"""
from .build import build_oracle, load_oracle  # noqa: F401
from .ref import (  # noqa: F401
    nnmf_ref,
    nnlm_ref,
    nnsvd_init_ref,
    icafast_ref,
    ica_init_ref,
    zero_replace_ref,
    run_nmf_ref,
    find_num_factors_ref,
    get_sample_coords_ref,
    factor_distance_ref,
    sammon_ref,
    get_factor_coords_ref,
    kl_div_ref,
)
