#!/usr/bin/env python3
"""Regenerate scenes/*.json (workload fixtures) from the reference checkout.

The five scene files under /root/reference/scenes are the *inputs* BASELINE.json's
configs name (they are data, not code).  /root/reference does not exist on the GPU
box, so the fixtures must travel with this repo.  This script re-emits every scene
as canonical one-line JSON: same schema (internal/scene/scene.go:9-158), same keys,
same numeric values (Python's float repr round-trips IEEE-754 binary64 exactly).

    python tools/import_scenes.py [/root/reference/scenes] [scenes]
"""
import json
import pathlib
import sys


def main() -> None:
    src = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/scenes")
    dst = pathlib.Path(sys.argv[2] if len(sys.argv) > 2 else pathlib.Path(__file__).resolve().parents[1] / "scenes")
    dst.mkdir(parents=True, exist_ok=True)
    for p in sorted(src.glob("*.json")):
        sc = json.loads(p.read_text())
        out = dst / p.name
        out.write_text(json.dumps(sc, separators=(",", ":"), ensure_ascii=False) + "\n")
        again = json.loads(out.read_text())
        assert again == sc, p
        print(f"{p.name}: {len(sc.get('objects', []))} objects, {len(sc.get('materials', []))} materials")


if __name__ == "__main__":
    main()
