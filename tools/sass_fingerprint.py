#!/usr/bin/env python
"""Fingerprints of the SASS of every kernel in quasi_b200/lib/libquasi_cuda.so.

    python tools/sass_fingerprint.py            # print them
    python tools/sass_fingerprint.py --update   # rewrite profiles/verified_sass.json from the current build
    python tools/sass_fingerprint.py --check    # exit 1 if a kernel listed there as verified has changed

profiles/verified_sass.json records, for the build the GPU parity suite last ran on, the md5 of each
kernel's instruction listing (cuobjdump -sass, encodings stripped). Editing a source file without
meaning to change the default kernels (a new build-time option, a test hook) must leave these
fingerprints alone: nvcc's output is sensitive even to statements that do nothing.
Refresh with --update only after `pytest -m gpu` has passed on the new build.
"""
import hashlib
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "quasi_b200" / "lib" / "libquasi_cuda.so"
RECORD = ROOT / "profiles" / "verified_sass.json"
# kernels that exist in the library but have not run on a GPU yet (none at the moment)
UNVERIFIED = re.compile(r"^$")


def fingerprints():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    out, name, lines = {}, None, []

    def flush():
        if name is not None:
            out[name] = {"md5": hashlib.md5("\n".join(lines).encode()).hexdigest(), "instructions": len(lines)}

    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            flush()
            mangled = m.group(1)
            demangled = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
            demangled = re.sub(r"\(anonymous namespace\)::|Q::cuda::|<unnamed>::", "", demangled)
            name, lines = re.sub(r"\(.*", "", demangled).replace("void ", ""), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", line)
        if m and name is not None:
            lines.append(m.group(2))
    flush()
    return out


def main():
    now = fingerprints()
    if "--update" in sys.argv:
        record = {"note": "see tools/sass_fingerprint.py; refresh only after the GPU suite has passed on this build",
                  "verified": {k: v for k, v in sorted(now.items()) if not UNVERIFIED.search(k)},
                  "not_yet_run_on_a_gpu": sorted(k for k in now if UNVERIFIED.search(k))}
        RECORD.write_text(json.dumps(record, indent=1) + "\n")
        print(f"wrote {RECORD}: {len(record['verified'])} verified kernels")
        return 0
    if "--check" in sys.argv:
        record = json.loads(RECORD.read_text())
        bad = [k for k, v in record["verified"].items() if now.get(k, {}).get("md5") != v["md5"]]
        for k in bad:
            print(f"CHANGED: {k}: {record['verified'][k]['instructions']} -> {now.get(k, {}).get('instructions')} instructions")
        print("all verified kernels unchanged" if not bad else f"{len(bad)} kernel(s) differ from the verified build")
        return 1 if bad else 0
    for k, v in sorted(now.items()):
        print(v["md5"], v["instructions"], k)
    return 0


if __name__ == "__main__":
    sys.exit(main())
