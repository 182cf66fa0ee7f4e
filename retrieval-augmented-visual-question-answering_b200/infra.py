"""Stand-ins for the handful of ``colbert.infra`` / ``colbert.data`` names the reference's call sites use
around the hot path, so that those call sites run verbatim where the reference package is not importable
(the GPU box), and the one function that addresses an index the way the reference does.

    Run, RunConfig, ColBERTConfig      third_party/ColBERT/colbert/infra/run.py:10-60,
                                       infra/config/{config,base_config,core_config,settings}.py
    Queries                            colbert/data/queries.py:13-82 (the ``data=`` dict form only)

What must agree with the reference (checked against the real package in tests/test_reference_callsites.py):

* ``Run()`` is a process-wide singleton holding a stack of ``RunConfig``; ``Run().context(cfg)`` pushes
  ``RunConfig.from_existing(Run().config, cfg)`` (run.py:50-61).
* a config remembers which of its fields were *assigned* (passed to the constructor — even as ``None`` —
  or ``configure``d); ``from_existing(*sources)`` overlays only assigned fields, later sources win
  (core_config.py:20-36, base_config.py:20-35).  The base ``Run()`` config has every field assigned
  (run.py:25-28), hence inside a ``Run().context(...)`` the ``root`` / ``experiment`` / ``index_root`` of
  the context always override those of a ``ColBERTConfig`` passed to ``Searcher``.
* ``index_root_ = index_root or <root>/<experiment>/indexes/`` (settings.py:50-52) and the Searcher opens
  ``os.path.join(index_root_, index)`` (searcher.py:26-30).

``resolve_index_path`` is duck-typed: it works with these classes and with the reference's own
(``colbert.infra.Run`` / ``ColBERTConfig``) when that package is loaded.
"""
from __future__ import annotations

import contextlib
import os
import sys
import time
from typing import Dict, Iterable, Optional


def _device_count() -> int:
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:      # pragma: no cover
        return 0


# field -> default, grouped like the reference's settings classes (settings.py:12-165)
_RUN_FIELDS: Dict[str, object] = {
    "overwrite": False,
    "root": os.path.join(os.getcwd(), "experiments"),       # evaluated once, at import (settings.py:20)
    "experiment": "default",
    "index_root": None,
    "name": time.strftime("%Y-%m/%d/%H.%M.%S"),
    "rank": 0,
    "nranks": 1,
    "amp": True,
    "total_visible_gpus": _device_count(),
    "gpus": _device_count(),
}
_COLBERT_ONLY_FIELDS: Dict[str, object] = {
    # ResourceSettings
    "checkpoint": None, "triples": None, "collection": None, "queries": None, "index_name": None,
    # DocSettings / QuerySettings
    "dim": 128, "doc_maxlen": 220, "mask_punctuation": True,
    "query_maxlen": 32, "attend_to_mask_tokens": False, "interaction": "colbert",
    # TrainingSettings
    "similarity": "cosine", "bsize": 32, "accumsteps": 1, "lr": 3e-06, "maxsteps": 500_000,
    "save_every": None, "resume": False, "warmup": None, "warmup_bert": None, "relu": False, "nway": 2,
    "use_ib_negatives": False, "reranker": False, "distillation_alpha": 1.0, "ignore_scores": False,
    # IndexingSettings / SearchSettings
    "index_path": None, "nbits": 1, "kmeans_niters": 4,
    "ncells": None, "centroid_score_threshold": None, "ndocs": None,
}


class _Config:
    """Field bag with 'assigned' tracking (core_config.py:20-62)."""
    FIELDS: Dict[str, object] = {}

    def __init__(self, **kw):
        unknown = set(kw) - set(self.FIELDS)
        if unknown:
            raise TypeError("%s got unexpected fields %s" % (type(self).__name__, sorted(unknown)))
        self.assigned = {}
        for name, default in self.FIELDS.items():
            value = kw.get(name)
            object.__setattr__(self, name, default if value is None else value)
            if name in kw:                      # an explicit None still counts as assigned
                self.assigned[name] = True

    # -- the reference's CoreConfig / BaseConfig surface ---------------------------------------------
    def assign_defaults(self):
        for name, default in self.FIELDS.items():
            setattr(self, name, default)
            self.assigned[name] = True

    def set(self, key, value, ignore_unrecognized=False):
        if key in self.FIELDS:
            setattr(self, key, value)
            self.assigned[key] = True
            return True
        if not ignore_unrecognized:
            raise Exception("Unrecognized key `%s` for %s" % (key, type(self)))

    def configure(self, ignore_unrecognized=True, **kw):
        return {k for k, v in kw.items() if not self.set(k, v, ignore_unrecognized)}

    def export(self):
        return {name: getattr(self, name) for name in self.FIELDS}

    @classmethod
    def from_existing(cls, *sources):
        kw = {}
        for src in sources:
            if src is None:
                continue
            kw.update({k: getattr(src, k) for k in getattr(src, "assigned", {}) if k in cls.FIELDS})
        return cls(**kw)

    @classmethod
    def from_deprecated_args(cls, args):
        obj = cls()
        return obj, obj.configure(ignore_unrecognized=True, **args)

    @classmethod
    def load_from_index(cls, index_path):
        """metadata.json, else plan.json: the ``config`` entry of either (base_config.py:71-87)."""
        import json
        for fn in ("metadata.json", "plan.json"):
            try:
                with open(os.path.join(index_path, fn)) as f:
                    args = json.load(f)
                return cls.from_deprecated_args(args.get("config", args))[0]
            except (OSError, ValueError):
                continue
        raise FileNotFoundError("no metadata.json / plan.json under %s" % index_path)

    # -- derived paths (settings.py:50-52, 148-150) ----------------------------------------------------
    @property
    def index_root_(self):
        return self.index_root or os.path.join(self.root, self.experiment, "indexes/")

    def __repr__(self):
        shown = ", ".join("%s=%r" % (k, getattr(self, k)) for k in self.assigned)
        return "%s(%s)" % (type(self).__name__, shown)


class RunConfig(_Config):
    FIELDS = dict(_RUN_FIELDS)


class ColBERTConfig(_Config):
    FIELDS = {**_RUN_FIELDS, **_COLBERT_ONLY_FIELDS}

    @property
    def index_path_(self):
        return self.index_path or os.path.join(self.index_root_, self.index_name)


class Run:
    """Singleton stack of run configurations (run.py:10-61)."""
    _instance = None

    def __new__(cls):
        if cls._instance is None:
            inst = super().__new__(cls)
            base = RunConfig()
            base.assign_defaults()
            inst.stack = [base]
            cls._instance = inst
        return cls._instance

    @property
    def config(self):
        return self.stack[-1]

    def __getattr__(self, name):
        # only reached for names that are not attributes of Run itself: forward to the active config
        stack = self.__dict__.get("stack")
        if stack and not name.startswith("__"):
            cfg = stack[-1]
            if name in cfg.FIELDS or isinstance(getattr(type(cfg), name, None), property):
                return getattr(cfg, name)
        raise AttributeError(name)

    @contextlib.contextmanager
    def context(self, runconfig: RunConfig, inherit_config: bool = True):
        if inherit_config:
            runconfig = RunConfig.from_existing(self.config, runconfig)
        self.stack.append(runconfig)
        try:
            yield
        finally:
            self.stack.pop()


class Queries:
    """``Queries(data={qid: text})`` (queries.py:13-45, 72-82): an ordered qid -> text mapping."""

    def __init__(self, path=None, data=None):
        if data is None:
            raise ValueError("only Queries(data={qid: text, ...}) is supported here (file loading is the "
                             "reference's data pipeline, out of scope)")
        assert isinstance(data, dict), type(data)
        self.path = path
        self.data = {qid: (c["question"] if isinstance(c, dict) else c) for qid, c in data.items()}

    def __len__(self):
        return len(self.data)

    def __iter__(self):
        return iter(self.data.items())

    def __getitem__(self, key):
        return self.data[key]

    def provenance(self):
        return self.path

    def keys(self):
        return self.data.keys()

    def values(self):
        return self.data.values()

    def items(self):
        return self.data.items()


# ---------------------------------------------------------------------------------------------------------
# index addressing
# ---------------------------------------------------------------------------------------------------------
def _reference_run_config():
    """``colbert.infra.run.Run().config`` when the reference package is loaded in this process."""
    mod = sys.modules.get("colbert.infra.run")
    if mod is None or not hasattr(mod, "Run"):
        return None
    try:
        return mod.Run().config
    except Exception:       # pragma: no cover
        return None


def active_run_config(config=None):
    """The ``Run().config`` the reference's Searcher would consult for ``config``: the reference's own
    singleton when ``config`` is one of its objects (or when only it has an open context), else ours."""
    ref = _reference_run_config()
    ours = Run().config
    if config is not None and type(config).__module__.startswith("colbert."):
        return ref if ref is not None else ours
    if isinstance(config, _Config) or ref is None:
        return ours
    ref_open = len(sys.modules["colbert.infra.run"].Run().stack) > 1
    return ref if (ref_open and len(Run().stack) == 1) else ours


def run_context_open() -> bool:
    """Is a ``Run().context(...)`` open — this package's or the reference's (when its package is loaded)?"""
    if len(Run().stack) > 1:
        return True
    mod = sys.modules.get("colbert.infra.run")
    return mod is not None and len(mod.Run().stack) > 1


def _overlay(sources: Iterable, keys=("index_root", "root", "experiment")) -> Dict[str, object]:
    """Assigned fields of ``sources`` overlaid in order (later wins) — ``from_existing`` restricted to the
    fields that address an index."""
    out: Dict[str, object] = {}
    for src in sources:
        if src is None:
            continue
        assigned = getattr(src, "assigned", None)
        for k in keys:
            if assigned is None:                # plain namespace / duck object: take what it has
                if getattr(src, k, None) is not None:
                    out[k] = getattr(src, k)
            elif k in assigned:
                out[k] = getattr(src, k)
    return out


def resolve_index_path(index: str, config=None, index_root: Optional[str] = None) -> str:
    """``os.path.join(ColBERTConfig.from_existing(config, Run().config).index_root_, index)``
    (colbert/searcher.py:26-30).  ``index_root=`` (an extension) bypasses the Run context; an absolute
    ``index`` wins over any root, as ``os.path.join`` makes it do in the reference."""
    if index_root is None:
        merged = _overlay([config, active_run_config(config)])
        index_root = merged.get("index_root") or os.path.join(
            merged.get("root", _RUN_FIELDS["root"]), merged.get("experiment", _RUN_FIELDS["experiment"]), "indexes/")
    return os.path.join(index_root, index)
