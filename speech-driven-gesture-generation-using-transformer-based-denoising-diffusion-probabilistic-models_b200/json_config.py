"""Attribute-style JSON config object, interface-compatible with the reference's
`utils/json_config.py:6-126` (`JsonConfig(path | dict | **kw)`, nested dicts become JsonConfig,
`Meta.name` injected from the file name, `.get`, `.to_dict()`, `.dump(path)`, `a + b` merge)."""
import json
import os


class JsonConfig(dict):
    Indent = 4

    def __init__(self, *argv, **kwargs):
        super().__init__()
        if argv and kwargs:
            raise AssertionError("[JsonConfig]: pass either one positional (json path or dict) or keyword items")
        if len(argv) > 1:
            raise AssertionError("[JsonConfig]: Need one positional parameters, found two.")
        src = argv[0] if argv else kwargs
        if isinstance(src, str):
            stem = os.path.splitext(os.path.basename(src))[0]
            with open(src, "r") as fh:
                src = json.load(fh)
            src.setdefault("Meta", {}).setdefault("name", stem)
        if not isinstance(src, dict):
            raise TypeError(f"[JsonConfig]: Do not support given input with type {type(src)}")
        for key, value in src.items():
            dict.__setitem__(self, key, JsonConfig(value) if isinstance(value, dict) else value)

    def __getattr__(self, attr):
        try:
            return dict.__getitem__(self, attr)
        except KeyError:
            raise KeyError(attr) from None

    def __getstate__(self):
        return self.__dict__

    def __setstate__(self, state):
        self.__dict__ = state

    def __add__(self, other):
        assert isinstance(other, JsonConfig)
        for key, value in other.items():
            if key not in self:
                dict.__setitem__(self, key, value)
            elif isinstance(value, JsonConfig):
                dict.__setitem__(self, key, self[key] + value)
            else:
                assert value == self[key], f"[JsonConfig]: Two config conflicts at`{key}`, {self[key]} != {value}"
        return self

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, JsonConfig) else v) for k, v in self.items() if not k.startswith("__")}

    def dump(self, json_path):
        with open(json_path, "w") as fh:
            json.dump(self.to_dict(), fh, indent=JsonConfig.Indent)

    def __str__(self):
        return json.dumps(self.to_dict(), indent=JsonConfig.Indent)


def normalize_model_config(config):
    """Accept both config schemas the reference ships.

    `configs/beat-ours.json` is flat (`Model.d_model`, `Model.Decoder.heads`, `Model.Generate`), which is what
    `models/model_creation.py:66,77-82,142` reads.  `configs/tedexp-ours.json` is a legacy nested form
    (`Model.Model.args.d_model`, `Model.Decoder.args.heads`, top-level `Generate`) that the shipped factory
    cannot load.  Returns `(model_params, d_pose_or_None, n_frames_or_None)` with model_params flat.
    `config` may be a whole-file config (has `Model`) or already the `Model` block.
    """
    root = config
    block = config["Model"] if ("Model" in config and "Decoder" not in config) else config
    d_pose = n_frames = None
    data = root.get("Data") if block is not root else None
    if data is not None:
        dargs = data.get("args", data)
        d_pose = dargs.get("pose_dim")
        n_frames = dargs.get("n_poses", dargs.get("pose_window_len"))
        if d_pose is None and "joints" in dargs and dargs.get("pose_representation") == "log_rot":
            d_pose = 3 * len(dargs["joints"])  # log-rotation: 3 numbers per joint (datasets/data_utils.py:101-107)
    if "Model" in block and isinstance(block["Model"], dict) and "args" in block["Model"]:
        flat = {"type": block["Model"]["type"], **block["Model"]["args"]}
        for part in ("Encoder", "Decoder", "Diffusion"):
            flat[part] = {"type": block[part]["type"], **block[part].get("args", {})}
        gen = block.get("Generate", root.get("Generate") if block is not root else None)
        if gen is not None:
            flat["Generate"] = dict(gen)
        block = JsonConfig(flat)
    elif not isinstance(block, JsonConfig):
        block = JsonConfig(dict(block))
    return block, d_pose, n_frames
