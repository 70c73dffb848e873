"""Record / compare per-kernel SASS fingerprints of libacez.so.

Kernels validated on the GPU must not change silently while unvalidated (opt-in) variants are added next to them in the
same translation units. After a validated GPU run:   python tools/sass_guard.py record profiles/validated_sass.json
Before a round ends without GPU time:                python tools/sass_guard.py check  profiles/validated_sass.json
`check` lists kernels whose SASS differs from the recorded fingerprint (addresses normalised), new kernels and missing ones.
"""
import hashlib
import json
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "acezero_b200" / "libacez.so"


def fingerprints(lib=LIB):
    out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    fp, cur, h = {}, None, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur:
                fp[cur] = h.hexdigest()
            cur, h = m.group(1), hashlib.sha1()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?)\s*/\*", line)
        if m and cur:
            h.update(re.sub(r"0x[0-9a-f]{4,}", "ADDR", m.group(1)).encode())
    if cur:
        fp[cur] = h.hexdigest()
    return fp


def main():
    mode, path = sys.argv[1], Path(sys.argv[2])
    fp = fingerprints()
    if mode == "record":
        path.write_text(json.dumps(fp, indent=1, sort_keys=True))
        print(f"recorded {len(fp)} kernels -> {path}")
        return 0
    old = json.loads(path.read_text())
    old = old.get("kernels", old)
    changed = [k for k in fp if k in old and old[k] != fp[k]]
    new = [k for k in fp if k not in old]
    missing = [k for k in old if k not in fp]
    for k in changed:
        print("CHANGED", k)
    for k in new:
        print("NEW    ", k)
    for k in missing:
        print("MISSING", k)
    print(f"{len(fp) - len(changed) - len(new)} unchanged, {len(changed)} changed, {len(new)} new, {len(missing)} missing")
    return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main())
