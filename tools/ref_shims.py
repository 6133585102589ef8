"""Import the UNMODIFIED reference: /root/reference in the build container, or the copy that
tools/install_reference.sh puts under baseline/_ref/ (git-ignored, shipped to the GPU box).

Test / baseline infrastructure only: used by tools/gen_tables.py and tools/make_golden.py to
produce committed numeric fixtures, by bench.py's CPU legs (`--impl reference`, `cpu_baseline`) and by
tests/test_dropin_reference.py.  Nothing in packppi_b200/ imports this file.

The reference needs ten third-party modules that are absent here.
They are only touched at import time or by code outside the sampling / proximal
path, so inert stand-ins registered in sys.modules are enough.
"""
import os
import sys
import types
import inspect

_HERE = os.path.dirname(os.path.abspath(__file__))
_INSTALLED = os.path.normpath(os.path.join(_HERE, "..", "baseline", "_ref"))  # tools/install_reference.sh


def _default_root():
    if os.environ.get("PACKPPI_REFERENCE"):
        return os.environ["PACKPPI_REFERENCE"]
    return "/root/reference" if os.path.isdir("/root/reference/src") else _INSTALLED


REFERENCE_ROOT = _default_root()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


class AttrDict(dict):
    """dict with attribute access (stands in for omegaconf.DictConfig)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    import torch

    if "pytorch_lightning" in sys.modules and getattr(sys.modules["pytorch_lightning"], "_pp_shim", False):
        return
    sys.dont_write_bytecode = True

    def rank_zero_only(fn):
        return fn

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            frame = inspect.currentframe().f_back
            loc = frame.f_locals
            hp = AttrDict()
            for key, val in loc.items():
                if key in ("self", "__class__"):
                    continue
                if key == "kwargs" and isinstance(val, dict):
                    hp.update(val)
                else:
                    hp[key] = val
            self.hparams = hp

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, *a, **k):
            pass

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return None

    pl = _module("pytorch_lightning", LightningModule=LightningModule, Callback=_Anything,
                 LightningDataModule=_Anything, Trainer=_Anything, _pp_shim=True)
    util = _module("pytorch_lightning.utilities", rank_zero_only=rank_zero_only)
    rz = _module("pytorch_lightning.utilities.rank_zero", rank_zero_only=rank_zero_only)
    lg = _module("pytorch_lightning.loggers", Logger=_Anything)
    pl.utilities, pl.loggers, util.rank_zero = util, lg, rz

    _module("omegaconf", DictConfig=AttrDict, OmegaConf=_Anything, open_dict=_Anything)
    hydra = _module("hydra", compose=None, initialize=None, main=lambda *a, **k: (lambda f: f))
    core = _module("hydra.core")
    hc = _module("hydra.core.hydra_config", HydraConfig=_Anything)
    hydra.core, core.hydra_config = core, hc
    _module("torchtyping", patch_typeguard=lambda: None, TensorType=_Anything)

    class MeanMetric(torch.nn.Module):
        def forward(self, *a, **k):
            return None

        def reset(self):
            pass

    _module("torchmetrics", MeanMetric=MeanMetric)
    _module("torch_scatter", scatter_add=None)
    _module("freesasa")
    bio = _module("Bio")
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
    import mini_biopdb  # the few Bio.PDB classes the reference's file readers touch
    biopdb = _module("Bio.PDB", PDBParser=mini_biopdb.PDBParser, NeighborSearch=mini_biopdb.NeighborSearch,
                     Selection=mini_biopdb.Selection)
    bio.PDB = biopdb
    _module("pyrootutils", setup_root=lambda *a, **k: None)

    class Data(AttrDict):
        def __init__(self, **kw):
            super().__init__(**kw)

        def to(self, device):
            for k, v in list(self.items()):
                if torch.is_tensor(v):
                    self[k] = v.to(device)
            return self

        def apply(self, fn):
            for k, v in list(self.items()):
                self[k] = fn(v)
            return self

        def clone(self):
            return Data(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.items()})

    tg = _module("torch_geometric")
    tgd = _module("torch_geometric.data", Data=Data)
    tg.data = tgd

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def import_reference(cache_folder="/tmp/packppi_ref_cache"):
    """Returns a namespace of the reference symbols on the hot path."""
    install()
    os.makedirs(cache_folder, exist_ok=True)
    import src.models.components.schedule as schedule

    if not getattr(schedule.SO2VESchedule, "_pp_wrapped", False):
        orig = schedule.SO2VESchedule.__init__

        def wrapped(self, *a, **k):
            k.setdefault("cache_folder", cache_folder)
            orig(self, *a, **k)

        schedule.SO2VESchedule.__init__ = wrapped
        schedule.SO2VESchedule._pp_wrapped = True

    import src.utils.residue_constants as rc
    import src.models.components as comp
    import src.models.components.encoder as encoder
    import src.models.components.mpnn as mpnn
    import src.models.components.layers as layers
    import src.models.components.clash as clash
    import src.models.components.optimize as optimize
    import src.utils.features as features
    import src.models.TorsionalDiffusion as td
    import src.datamodules.components.complex_dataset as cds
    import src.datamodules.components.helper as helper
    from torch_geometric.data import Data

    return types.SimpleNamespace(rc=rc, comp=comp, encoder=encoder, mpnn=mpnn, layers=layers, clash=clash,
                                 optimize=optimize, features=features, td=td, schedule=schedule, cds=cds,
                                 helper=helper, Data=Data, AttrDict=AttrDict)


REF_CFG = dict(
    encoder_cfg=dict(node_in=35, edge_in=468, node_features=128, edge_features=128,
                     time_embedding_type="sinusoidal", time_embedding_dim=16, num_positional_embeddings=16,
                     num_rbf=16, top_k=32, af2_relpos=True),
    model_cfg=dict(hidden_dim=128, num_mpnn_layers=3, n_points=8, dropout=0.1, act="relu",
                   position_scale=1.0, use_ipmp=True, k_neighbors=32),
    sample_cfg=dict(eval_epochs=1, sample_during_training=True, annealed_temp=3, mode="ode", use_proximal=True,
                    violation_tolerance_factor=12., clash_overlap_tolerance=0.5, lamda=1., num_steps=50),
)


def build_reference_model(ref):
    """TDiffusionModule from the reference constructor with the yaml values (configs/model/*)."""
    cfg = {k: AttrDict(v) for k, v in REF_CFG.items()}
    model = ref.td.TDiffusionModule(optimizer=None, scheduler=None, **cfg)
    return model.eval()
