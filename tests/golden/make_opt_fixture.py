"""Build `opt` with the reference's own options.py from its shipped YAMLs and commit it as JSON.

Run in the build container only (needs /root/reference):  python tests/golden/make_opt_fixture.py
The reference's `options.set` (options.py:59-70) = load_options (YAML + `_parent_` chain) -> override_options (command
line) -> process_options (seed, output directory, device, H/W). The first two are called unmodified; of
process_options only the three assignments the field layer can see are restated (device, H, W) -- the rest creates an
output directory under the (read-only) reference tree.
"""
import contextlib
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

YAMLS = {"bat_blender_VM_MLP": "bat", "bat_llff_VM_MLP": "bat", "bat_blender_VM": "bat"}


def build(yaml_name, model):
    ref_loader._install_shims()
    root = ref_loader.REFERENCE_ROOT
    if root not in sys.path:
        sys.path.insert(0, root)
    cwd = os.getcwd()
    os.chdir(root)                       # options.py opens "options/<name>.yaml" relative to the working directory
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import options as ref_options
            import util as ref_util
            from easydict import EasyDict as edict
            opt = ref_options.load_options(f"options/{yaml_name}.yaml")
            opt = ref_options.override_options(opt, edict(model=model, yaml=yaml_name), key_stack=[], safe_check=False)
            opt.device = "cuda:0"
            opt.H, opt.W = map(int, opt.data.image_size)
            return ref_util.to_dict(opt)
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    for y, mdl in YAMLS.items():
        d = build(y, mdl)
        with open(os.path.join(HERE, f"opt_{y}.json"), "w") as f:
            json.dump(d, f, indent=1, sort_keys=True, default=str)
        print(y, "->", len(json.dumps(d)), "bytes;", "arch.tensorf.model =", d["arch"]["tensorf"]["model"])
