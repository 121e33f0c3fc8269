"""Drop-in launcher: run the reference's UNCHANGED host scripts on top of hortimapping_b200.

    python -m hortimapping_b200.dropin /path/to/HortiMapping/run_shape_completion_challenge.py -c configs/...yaml

`install(reference_root)` puts the reference root on sys.path (so `wild_completion.utils`,
`wild_completion.opt_visualizer`, `dataloader`, `metrics_3d` keep resolving to the reference's own code)
and pre-registers replacement modules under the three module paths the host scripts import the hot path
from (test_wild_completion.py:15-21, run_shape_completion_challenge.py:14-22):

    wild_completion.optimizer.Optimizer            -> hortimapping_b200.optimizer.Optimizer
    wild_completion.mesher.MeshExtractor           -> hortimapping_b200.mesher.MeshExtractor
    deepsdf.deep_sdf.workspace.{config_decoder,load_latent_vectors} -> hortimapping_b200.decoder
    metrics_3d.chamfer_distance.ChamferDistance, metrics_3d.precision_recall.PrecisionRecall -> hortimapping_b200.metrics
    wild_completion.utils.get_render_data / get_rays (attributes of the reference's own module) -> hortimapping_b200.render_data
    wild_completion.utils.clean_mesh / clean_pcd / get_pose_init (utils.py:389-459)             -> hortimapping_b200.preprocess
"""
from __future__ import annotations

import importlib
import os
import runpy
import sys
import types


def install(reference_root: str | None = None) -> None:
    from . import decoder as _decoder
    from . import mesher as _mesher
    from . import optimizer as _optimizer
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)

    def module(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    # deepsdf.deep_sdf.workspace (workspace.py:82-114,203-225); wild_completion/utils.py:20 imports from it too
    for pkg in ("deepsdf", "deepsdf.deep_sdf"):
        if pkg not in sys.modules:
            p = module(pkg)
            p.__path__ = []       # mark as package
    ws = module("deepsdf.deep_sdf.workspace", config_decoder=_decoder.config_decoder, load_latent_vectors=_decoder.load_latent_vectors)
    sys.modules["deepsdf.deep_sdf"].workspace = ws
    sys.modules["deepsdf"].deep_sdf = sys.modules["deepsdf.deep_sdf"]
    # the reference's own wild_completion package stays importable for utils / opt_visualizer
    try:
        wc = importlib.import_module("wild_completion")
    except ImportError:
        wc = module("wild_completion")
        wc.__path__ = [os.path.join(reference_root, "wild_completion")] if reference_root else []
    opt_mod = module("wild_completion.optimizer", Optimizer=_optimizer.Optimizer)
    mesh_mod = module("wild_completion.mesher", MeshExtractor=_mesher.MeshExtractor)
    loss_mod = module("wild_completion.loss", compute_sdf_loss=_optimizer.compute_sdf_loss, compute_render_loss=_optimizer.compute_render_loss)
    wc.optimizer, wc.mesher, wc.loss = opt_mod, mesh_mod, loss_mod
    # evaluation metrics (metrics_3d/chamfer_distance.py, precision_recall.py): same classes, GPU nearest neighbours
    from . import metrics as _metrics
    try:
        m3 = importlib.import_module("metrics_3d")
    except Exception:
        m3 = module("metrics_3d", MESHTYPE=6, TETRATYPE=10, PCDTYPE=1)
        m3.__path__ = [os.path.join(reference_root, "metrics_3d")] if reference_root else []
    m3.chamfer_distance = module("metrics_3d.chamfer_distance", ChamferDistance=_metrics.ChamferDistance)
    m3.precision_recall = module("metrics_3d.precision_recall", PrecisionRecall=_metrics.PrecisionRecall)
    # the step before the hot path: wild_completion.utils stays the reference's own module, only get_render_data / get_rays
    # (utils.py:23-109) are rebound to the device versions (same signature, same dict, same np.random draws)
    try:
        from . import render_data as _rd
        wu = importlib.import_module("wild_completion.utils")
        wu.get_render_data, wu.get_rays = _rd.get_render_data, _rd.get_rays
        # ... and the point-cloud preparation of utils.py:389-459 (DBSCAN cleaning, pose initialisation) to the device versions
        from . import preprocess as _pp
        wu.clean_mesh, wu.clean_pcd, wu.get_pose_init = _pp.clean_mesh, _pp.clean_pcd, _pp.get_pose_init
    except Exception:
        pass      # the reference's utils needs open3d / skimage; without them the host scripts cannot run anyway


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    script = os.path.abspath(argv[0])
    install(os.path.dirname(script))
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
