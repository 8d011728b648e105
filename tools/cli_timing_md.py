"""profiles/<tag>_cli_timing.md from the JSON lines of tools/cli_bench.py:  python tools/cli_timing_md.py out.md a.json b.json"""
import json
import sys


def last_json(path):
    for line in reversed(open(path).read().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise SystemExit(f"no JSON line in {path}")


rows = [last_json(p) for p in sys.argv[2:]]
out = ["# CLI, files in -> files out (`python tools/cli_bench.py`: wall clock of `cli.main`, best of the warm repetitions)", "",
       "| job | `main` (batched pipeline, `cli.run_fast`) | `--reference_loop` (per-image `preprocess` / `run_batch`) |", "|---|---|---|"]
for r in rows:
    job = f"{r['frames']} PNGs of {r['size']}^2, --batch {r['batch']}, --img_size {r['img_size']}"
    f, l = r["fast"], r["reference_loop"]
    out.append(f"| {job} | {f['seconds_best']:.3f} s = {f['images_per_s']:.1f} img/s | {l['seconds_best']:.3f} s = {l['images_per_s']:.1f} img/s |")
out += ["",
        "cProfile of the 64 x 1024^2 job before the round's last two host changes (`tools/cli_profile.py`, 2.43 s): pandas `to_csv` 1.88 s",
        "(float -> text: 64 per-image files 1.0 s + `all_droplets.csv` 0.9 s for 195 k rows), `cv2.imwrite` 0.6 s (writer pool), PNG decode",
        "0.36 s (decoder pool), the GPU pipeline 0.08 s.  The 8 x 256^2 job (0.38 s): `UNetDC.__init__` random init 0.12 s, weight packing",
        "0.09 s, `torch.load` 0.05 s, pipeline 0.02 s.  Both are host-bound: checkpoint loading and text formatting, not the device path --",
        "hence the meta-device model construction in `cli.load_model` and `all_droplets.csv` written from the per-image CSV bodies."]
open(sys.argv[1], "w").write("\n".join(out) + "\n")
print("\n".join(out))
