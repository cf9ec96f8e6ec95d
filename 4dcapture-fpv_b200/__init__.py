"""4dcapture-fpv_b200 -- B200-native hot path of 4DCapture-FPV's global_optimization stage.

The directory name is not a Python identifier; import it with
    import importlib; fpv = importlib.import_module("4dcapture-fpv_b200")
(tests/conftest.py and bench.py also alias it as `fpv_b200` in sys.modules).

Public surface (each mirrors a reference signature, see the module docstrings):
    distChamfer(a, b), chamferDist(), pairwise_dist,
    NN_loss, body_to_scene, scene_to_body,
    scene_to_body_sum, SearchOptions, SearchState         chamfer.py
    create(...), SMPLXB200                                body_model.py
    verts_transform, body2world, contact_robust_loss,
    second_diff_l1, first_diff_l1                         residuals.py
    convert_to_3D_rot, convert_to_6D_rot,
    VPoserDecoderB200, cal_dctloss                        prior.py   
    distChamferSharded, allreduce_grads, shard_range      sharded.py
    Mailbox (peer-memory exchange of the sharded step)    p2p.py
    io_formats (pickles, camerapose.txt, PLY)             io_formats.py  (host-side)
    FitProblem                                            fit.py
Everything computes on an sm_100 GPU through libfpv_b200.so; there is no CPU or eager fallback.
"""
from . import _lib  # noqa: F401
from .body_model import SMPLXB200, SMPLXOutput, create, load_smplx_npz  # noqa: F401
from .chamfer import (NN_loss, SearchOptions, SearchState, body_to_scene, chamferDist, distChamfer,  # noqa: F401
                      fit_chamfer_terms, nn_search, pack_planes, pairwise_dist, scene_to_body, scene_to_body_sum,
                      unpack_keys)
from .fit import FitProblem, LOSS_WEIGHTS  # noqa: F401
from .residuals import (body2world, contact_robust_loss, first_diff_l1, second_diff_l1,  # noqa: F401
                        verts_transform)
from .prior import (VPoserDecoderB200, aa_to_rot6d, body_params_encapsulate_batch, cal_dctloss, front_end_split,  # noqa: F401
                    convert_to_3D_rot, convert_to_6D_rot, dct_basis, make_vposer_weights, rot6d_to_aa)
from .sharded import allreduce_grads, combine_keys, distChamferSharded, shard_range  # noqa: F401
from . import io_formats, p2p, sharded, spatial, synthetic  # noqa: F401

__version__ = "0.1.0"
