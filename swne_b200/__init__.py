"""swne_b200 -- B200-native (sm_100a) NMF hot path of SWNE: RunNMF -> NNLM::nnmf, seeding and the
get_factor_coords / get_sample_coords consumers, behind the C-ABI in include/swne_nmf.h."""
from .nmf import (FindNumFactors, NMFHandle, ProjectFeatures, ProjectSamples, RunNMF, ica_init, kl_div, nnmf,  # noqa: F401
                  nnsvd_init, zero_replace)
from .embed import EmbedSWNE, get_factor_coords, get_sample_coords, sammon  # noqa: F401
