"""HyperPyYAML-subset loader for the reference's ``hparams/*.yaml`` (the layer shell).

The reference recipes describe the whole experiment in HyperPyYAML (``train_CTC.py:842``: ``load_hyperpyyaml(fin,
overrides)``): plain YAML plus the tags ``!ref`` (reference / interpolation / arithmetic), ``!new:`` (instantiate a class),
``!name:`` (a callable with bound keyword arguments), ``!apply:`` (call a function), ``!PLACEHOLDER``.  This module reads the
same files without speechbrain / hyperpyyaml installed:

* every tag is parsed and every ``!ref`` resolved, so the scalar hyper-parameters (``d_model``, ``n_fft``, ``max_grad_norm``,
  ``lr_model`` ...) come out exactly as the reference computes them;
* ``!new:`` / ``!name:`` / ``!apply:`` of objects ON the hot path are built from this package (table ``REGISTRY``: Fbank,
  InputNormalization, the conv front-end, the TransformerASR wrapper, Linear, the Noam schedule) or from ``torch.*`` itself;
* everything else (augmentation, tokenizers, loggers, checkpointing, beam search: out of scope per SURVEY.md section 8) becomes
  an ``Unresolved`` record that keeps the class path and the resolved keyword arguments and raises if it is called.

``model_config(hp)`` extracts the model block in the form ``encoder.CONFIGS`` uses and ``build_model_from_hparams(hp)``
constructs the B200-native model (``ConMambaCTC`` / ``ConMambaS2S``) from it, so a reference YAML selects the same network,
feature front-end and optimizer settings here as it does there.
"""
import ast
import functools
import importlib
import operator
import re

import yaml

__all__ = ["load_hparams", "model_config", "build_model_from_hparams", "optimizer_config", "Unresolved", "Placeholder"]


class Placeholder:
    """``!PLACEHOLDER``: a value the caller must supply through ``overrides``."""

    def __repr__(self):
        return "<!PLACEHOLDER>"


class Unresolved:
    """A ``!new:`` / ``!name:`` / ``!apply:`` target outside the hot path: kept as data, never executed."""

    def __init__(self, kind, path, args, kwargs):
        self.kind, self.path, self.args, self.kwargs = kind, path, args, kwargs

    def __call__(self, *a, **k):
        raise NotImplementedError("%s:%s is outside the B200 hot path (SURVEY.md section 8: out of scope)" % (self.kind, self.path))

    def __repr__(self):
        return "<Unresolved !%s:%s>" % (self.kind, self.path)


class Spec:
    """What the reference would have instantiated, with the resolved arguments (``TransformerASR``, ``ConvolutionFrontEnd`` ...):
    ``build_model_from_hparams`` turns the specs of a file into one model."""

    def __init__(self, path, **kwargs):
        self.path, self.kwargs = path, dict(kwargs)

    def __repr__(self):
        return "<Spec %s %s>" % (self.path, sorted(self.kwargs))


class NoamSchedule:
    """speechbrain.nnet.schedulers.NoamScheduler as the recipes use it: ``lr = schedule(step)`` (trainer.noam_lr)."""

    def __init__(self, lr_initial, n_warmup_steps, model_size=None):
        self.lr_initial, self.n_warmup_steps, self.model_size = lr_initial, n_warmup_steps, model_size

    def __call__(self, step):
        from .trainer import noam_lr
        return noam_lr(self.lr_initial, self.n_warmup_steps, step)


def _fbank(**kw):
    from .fbank import Fbank
    return Fbank(**kw)


def _input_norm(**kw):
    from .encoder import InputNormalization
    return InputNormalization()


def _linear(input_size, n_neurons, bias=True, **kw):
    import torch.nn as nn
    return nn.Linear(input_size, n_neurons, bias=bias)


# reference class path -> constructor here (hot path only)
REGISTRY = {
    "speechbrain.lobes.features.Fbank": _fbank,
    "speechbrain.processing.features.InputNormalization": _input_norm,
    "speechbrain.nnet.linear.Linear": _linear,
    "speechbrain.nnet.schedulers.NoamScheduler": NoamSchedule,
    "speechbrain.lobes.models.convolution.ConvolutionFrontEnd": functools.partial(Spec, "ConvolutionFrontEnd"),
    "modules.TransformerASR.TransformerASR": functools.partial(Spec, "TransformerASR"),
}


def _import_path(path):
    mod, _, name = path.rpartition(".")
    return getattr(importlib.import_module(mod), name)


def _resolve_callable(kind, path):
    if path in REGISTRY:
        return REGISTRY[path]
    if path.startswith("torch."):
        return _import_path(path)
    return None


# ---- parsing: YAML nodes -> a tree of plain values and tagged records -------------------------------------------------
class _Tagged:
    def __init__(self, kind, path, value):
        self.kind, self.path, self.value = kind, path, value


class _Loader(yaml.SafeLoader):
    pass


def _construct_any(loader, node):
    if isinstance(node, yaml.MappingNode):
        return loader.construct_mapping(node, deep=True)
    if isinstance(node, yaml.SequenceNode):
        return loader.construct_sequence(node, deep=True)
    return loader.construct_scalar(node)


def _multi(kind):
    def ctor(loader, suffix, node):
        val = None if (isinstance(node, yaml.ScalarNode) and node.value == "") else _construct_any(loader, node)
        return _Tagged(kind, suffix, val)
    return ctor


_Loader.add_multi_constructor("!new:", _multi("new"))
_Loader.add_multi_constructor("!name:", _multi("name"))
_Loader.add_multi_constructor("!apply:", _multi("apply"))
_Loader.add_constructor("!ref", lambda loader, node: _Tagged("ref", None, loader.construct_scalar(node)))
_Loader.add_constructor("!PLACEHOLDER", lambda loader, node: Placeholder())
_Loader.add_constructor("!tuple", lambda loader, node: tuple(loader.construct_sequence(node, deep=True)))


# ---- resolution ---------------------------------------------------------------------------------------------------------
_REF = re.compile(r"<([A-Za-z_][A-Za-z0-9_.\[\]]*)>")
_OPS = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv,
        ast.FloorDiv: operator.floordiv, ast.Pow: operator.pow, ast.Mod: operator.mod, ast.USub: operator.neg}


def _arith(expr):
    """Value of an arithmetic expression over numbers (what HyperPyYAML evaluates inside ``!ref``), or None."""
    try:
        tree = ast.parse(expr.strip(), mode="eval").body
    except SyntaxError:
        return None

    def ev(n):
        if isinstance(n, ast.Constant) and isinstance(n.value, (int, float)) and not isinstance(n.value, bool):
            return n.value
        if isinstance(n, ast.BinOp) and type(n.op) in _OPS:
            return _OPS[type(n.op)](ev(n.left), ev(n.right))
        if isinstance(n, ast.UnaryOp) and type(n.op) in _OPS:
            return _OPS[type(n.op)](ev(n.operand))
        raise ValueError
    try:
        return ev(tree)
    except (ValueError, ZeroDivisionError):
        return None


def _literal(value):
    """YAML would read ``(8, 10, 80)`` as a string; HyperPyYAML turns tuple-looking strings into tuples."""
    if isinstance(value, str):
        s = value.strip()
        if len(s) >= 2 and s[0] == "(" and s[-1] == ")":
            try:
                v = ast.literal_eval(s)
                if isinstance(v, tuple):
                    return v
            except (ValueError, SyntaxError):
                pass
    return value


class _Resolver:
    def __init__(self, tree):
        self.tree = tree
        self.cache = {}
        self.active = set()

    def lookup(self, dotted):
        node = self.tree
        trail = []
        for part in dotted.split("."):
            m = re.fullmatch(r"([A-Za-z_][A-Za-z0-9_]*)((\[\d+\])*)", part)
            if m is None:
                raise KeyError(dotted)
            trail.append(m.group(1))
            key = ".".join(trail)
            if isinstance(node, dict) and m.group(1) in node:
                # resolve (and memoise) top-level keys so that objects referenced twice are ONE object, as in HyperPyYAML
                if len(trail) == 1:
                    node = self.top(m.group(1))
                else:
                    node = self.resolve(node[m.group(1)]) if isinstance(node[m.group(1)], (_Tagged,)) else node[m.group(1)]
            else:
                raise KeyError("!ref <%s>: no such key (%s)" % (dotted, key))
            for idx in re.findall(r"\[(\d+)\]", m.group(2)):
                node = node[int(idx)]
        return node

    def top(self, key):
        if key in self.cache:
            return self.cache[key]
        if key in self.active:
            raise ValueError("circular !ref through <%s>" % key)
        self.active.add(key)
        val = self.resolve(self.tree[key])
        self.active.discard(key)
        self.cache[key] = val
        return val

    def resolve(self, v):
        if isinstance(v, _Tagged):
            if v.kind == "ref":
                return self.ref(v.value)
            return self.build(v)
        if isinstance(v, dict):
            return {k: self.resolve(x) for k, x in v.items()}
        if isinstance(v, list):
            return [self.resolve(x) for x in v]
        return _literal(v)

    def ref(self, text):
        text = text.strip()
        whole = _REF.fullmatch(text)
        if whole is not None:                       # "!ref <key>": the object itself
            return self.lookup(whole.group(1))
        parts = {}

        def sub(m):
            val = self.lookup(m.group(1))
            parts[m.group(0)] = val
            return repr(val) if isinstance(val, (int, float)) and not isinstance(val, bool) else str(val)
        flat = _REF.sub(sub, text)
        if parts and all(isinstance(x, (int, float)) and not isinstance(x, bool) for x in parts.values()):
            val = _arith(flat)                      # "!ref <a> * 2": arithmetic on numbers
            if val is not None:
                return val
        return flat                                 # "!ref <folder>/save": string interpolation

    def build(self, t):
        val = self.resolve(t.value) if t.value is not None else None
        args, kwargs = [], {}
        if isinstance(val, dict):
            kwargs = val
        elif isinstance(val, list):
            args = val
        elif val is not None:
            args = [val]
        fn = _resolve_callable(t.kind, t.path)
        if fn is None:
            return Unresolved(t.kind, t.path, args, kwargs)
        if t.kind == "name":
            return functools.partial(fn, *args, **kwargs) if (args or kwargs) else fn
        if t.path.startswith("torch.") and t.kind == "new":
            try:
                return fn(*args, **kwargs)
            except (TypeError, ValueError):         # e.g. a ModuleList over Spec records: keep it as data
                return Unresolved(t.kind, t.path, args, kwargs)
        return fn(*args, **kwargs)                  # !new / !apply


def load_hparams(source, overrides=None):
    """``source``: YAML text or an open file of a reference ``hparams/*.yaml``; ``overrides``: dict of top-level replacements
    (what the recipes pass on the command line, e.g. ``data_folder``).  Returns the resolved dict."""
    text = source.read() if hasattr(source, "read") else source
    tree = yaml.load(text, Loader=_Loader) or {}
    if not isinstance(tree, dict):
        raise ValueError("a hyperparameter file is a mapping at the top level")
    for k, v in (overrides or {}).items():
        tree[k] = v
    r = _Resolver(tree)
    return {k: r.top(k) for k in tree}


def model_config(hp):
    """The model block of a resolved reference YAML in the form of ``encoder.CONFIGS`` entries."""
    tr = hp.get("Transformer")
    if not isinstance(tr, Spec):
        raise ValueError("no `Transformer: !new:modules.TransformerASR.TransformerASR` in this file")
    k = tr.kwargs
    if k.get("encoder_module") != "conmamba":
        raise NotImplementedError("encoder_module=%r: only the ConMamba encoder is on the B200 path (SURVEY.md section 8)"
                                  % k.get("encoder_module"))
    mc = dict(k.get("mamba_config") or {})
    cfg = dict(d_model=k["d_model"], d_ffn=k["d_ffn"], num_layers=k["num_encoder_layers"],
               n_fft=hp.get("n_fft", 400), win_length=hp.get("win_length", 25), n_mels=hp.get("n_mels", 80),
               output_neurons=k.get("tgt_vocab", hp.get("output_neurons")), dropout=k.get("dropout", 0.1),
               d_state=mc.get("d_state", 16), expand=mc.get("expand", 2), d_conv=mc.get("d_conv", 4),
               bidirectional=mc.get("bidirectional", True), seed=hp.get("seed"))
    ndec = k.get("num_decoder_layers", 0) or 0
    if ndec > 0:
        if k.get("decoder_module") != "mamba":
            raise NotImplementedError("decoder_module=%r: only the Mamba decoder is on the B200 path" % k.get("decoder_module"))
        cfg["num_decoder_layers"] = ndec
        cfg.pop("bidirectional")                    # ConMambaS2S fixes the encoder to bidirectional
    cnn = hp.get("CNN")
    if isinstance(cnn, Spec):
        ck = cnn.kwargs
        if tuple(ck.get("out_channels", (64, 32))) != (64, 32) or tuple(ck.get("strides", (2, 2))) != (2, 2) or \
                tuple(ck.get("kernel_sizes", (3, 3))) != (3, 3) or ck.get("num_layers_per_block", 1) != 1:
            raise NotImplementedError("ConvolutionFrontEnd variant outside the reference recipes: %r" % (ck,))
    return cfg


def optimizer_config(hp):
    """Arguments of ``trainer.TrainStep`` from a resolved reference YAML (AdamW + Noam + clipping)."""
    noam = hp.get("noam_annealing")
    out = dict(lr=hp.get("lr_model", hp.get("lr_adam", 1e-3)), weight_decay=hp.get("weight_decay", 0.0),
               max_grad_norm=hp.get("max_grad_norm", 5.0))
    if isinstance(noam, NoamSchedule):
        out["lr"], out["n_warmup_steps"] = noam.lr_initial, noam.n_warmup_steps
    opt = hp.get("model_opt_class", hp.get("Adam"))
    if isinstance(opt, functools.partial):
        kw = opt.keywords
        out["lr"] = kw.get("lr", out["lr"])
        if "betas" in kw:
            out["betas"] = tuple(kw["betas"])
        if "eps" in kw:
            out["eps"] = kw["eps"]
        out["weight_decay"] = kw.get("weight_decay", out["weight_decay"])
    return out


def build_model_from_hparams(hp):
    """The B200-native model a reference YAML describes (ConMambaCTC, or ConMambaS2S when it has a Mamba decoder)."""
    import torch
    from .encoder import ConMambaCTC, ConMambaS2S
    cfg = model_config(hp)
    seed = cfg.pop("seed", None)
    if seed is not None:
        torch.manual_seed(seed)
    if "num_decoder_layers" in cfg:
        return ConMambaS2S(**cfg)
    return ConMambaCTC(**cfg)
