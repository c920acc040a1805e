#!/usr/bin/env python3
"""Generates tests/golden/*.json from the reference checkout (run in the build container only;
/root/reference does not exist on the GPU box).

scene_save.json: sha256 of the reference scene files that were written by the reference's own scene.Save
(internal/scene/io.go:25-38) — the only reference-produced artefacts in the tree.  They pin our encoder's
layout and float formatting.
"""
import hashlib
import json
import pathlib

REF = pathlib.Path("/root/reference/scenes")
HERE = pathlib.Path(__file__).resolve().parent


def main():
    out = {}
    for name in ["metal_glass_room.json"]:
        data = (REF / name).read_bytes()
        out[name] = {"sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data)}
    (HERE / "scene_save.json").write_text(json.dumps(out, indent=1) + "\n")
    print(out)


if __name__ == "__main__":
    main()
