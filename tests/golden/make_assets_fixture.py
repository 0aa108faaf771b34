"""Generate tests/golden/pco_assets.json from the reference's golden .pco assets.

Run once in the authoring container (where /root/reference exists):
    python tests/golden/make_assets_fixture.py
The assets are the reference's backward-compatibility decoder fixtures
(pco/assets/*.pco, exercised by pco/src/tests/compatibility.rs:70-303).  They are
stored hex-encoded with a sha256 so the GPU box (which has no /root/reference)
can still run the golden tests.
"""
import hashlib
import json
import pathlib

ASSET_DIR = pathlib.Path("/root/reference/pco/assets")
OUT = pathlib.Path(__file__).with_name("pco_assets.json")


def main():
    out = {}
    for p in sorted(ASSET_DIR.glob("*.pco")):
        data = p.read_bytes()
        out[p.stem] = {"hex": data.hex(), "sha256": hashlib.sha256(data).hexdigest(), "len": len(data)}
    OUT.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print(f"wrote {OUT} ({len(out)} assets)")


if __name__ == "__main__":
    main()
