"""vvc_intra_b200 -- B200-native intra cost-evaluation engine behind the IntraSearch::estIntraPredLumaQT /
EncCu::xCompressCU call surface of the VTM 6.1 fork llsurreal919/Reduce-Complexity-for-intra-coding-of-VVC.

The compute path is hand-written CUDA for sm_100a behind the C ABI of include/vvc_intra_b200.h
(libvvc_intra_b200.so, built in-tree by __graft_entry__.build()).  This package is the thin host-side
mirror of that ABI plus the host logic of the exhaustive candidate sweep; it has no CPU fallback."""
from .engine import (IntraCostEngine, EngineError, TU_JOB_DTYPE, TU_SRC_DTYPE, TU_RESULT_DTYPE, OPT_DEP_QUANT, TU_QUANT, TU_DEPQUANT, TU_RDOQ_TS, TU_TS_ALLOWED, TU_MTS_ALLOWED, TU_RATE, BIN_MODEL_DTYPE, CTX_STATES_DTYPE, DQ_RATES_DTYPE, FEAT_CU_DTYPE, FEAT_JOB_DTYPE, FEAT_RESULT_DTYPE, VISIT_DTYPE, RESULT_DTYPE, DETAIL_DTYPE, NUM_SLOTS, SLOT_MRL1, SLOT_MRL3, SLOT_MIP,
                     SAT_NONE, RECT_DTYPE, BRIEF_DTYPE, CU_AUTO_DTYPE, AUTO_FINAL, AUTO_REGULAR, OPT_YIELD_SYNC, OPT_TRUSTED_VISITS, library_path, ISP_PART_DTYPE, ISP_MODE_DTYPE, ISP_HOR, ISP_VER, TR_DCT2, TR_DCT8, TR_DST7)
from .features import feature_gate, select_feature_neighbours, feature_job, classifier_decision
from .partition import enumerate_root_candidates, frame_candidates, candidate_availability, build_sweep_visits

__all__ = ['IntraCostEngine', 'EngineError', 'VISIT_DTYPE', 'RESULT_DTYPE', 'DETAIL_DTYPE', 'NUM_SLOTS', 'SLOT_MRL1', 'SLOT_MRL3',
           'SLOT_MIP', 'SAT_NONE', 'library_path', 'enumerate_root_candidates', 'frame_candidates',
           'candidate_availability', 'build_sweep_visits']
from . import shard
from .tu_sweep import build_tu_sweep, build_tu_jobs_from_lists, slots_of_modes, default_dq_rates, default_ctx_states, lambda_for_qp
from . import assemble, hls, training_set          # frame_parallel (a driver with a CLI) is imported on demand
