"""Fingerprint of every kernel's SASS in a built library (no GPU needed):
    python tools/sass_fingerprint.py word2bits_b200/libw2b.so [--save F | --compare F] [--filter SUBSTR]
Used to prove that a source change left the machine code of the production kernels untouched (e.g. when an
experimental variant is added behind a template flag): instruction text with encodings stripped and branch
labels renumbered in order of appearance, hashed per kernel.  Template arguments are part of the key, so a
kernel that gains a template parameter is compared under --map OLD=NEW substitutions of its mangled name."""
import hashlib
import json
import re
import subprocess
import sys


def fingerprints(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out = {}
    for part in re.split(r"\n\s*Function : ", txt)[1:]:
        name, body = part.split("\n", 1)
        lines, labels = [], {}
        for l in body.split("\n"):
            l = re.sub(r"/\*[0-9a-fx]+\*/", "", l).strip()
            if not l or l.startswith(".headerflags") or l.startswith("Fatbin") or l.startswith("===") or \
                    l.startswith("arch =") or l.startswith("code version") or l.startswith("host =") or \
                    l.startswith("compile_size") or l.startswith("producer") or l.startswith("identifier"):
                continue
            l = re.sub(r"\.L_x_\d+", lambda m: labels.setdefault(m.group(0), ".L%d" % len(labels)), l)
            lines.append(l)
        out[name.strip()] = [len(lines), hashlib.md5("\n".join(lines).encode()).hexdigest()]
    return out


def main():
    lib = sys.argv[1]
    args = sys.argv[2:]
    flt = args[args.index("--filter") + 1] if "--filter" in args else ""
    maps = [a.split("=", 1) for i, a in enumerate(args) if i and args[i - 1] == "--map"]
    fp = {k: v for k, v in fingerprints(lib).items() if flt in k}
    if "--save" in args:
        json.dump(fp, open(args[args.index("--save") + 1], "w"), indent=0, sort_keys=True)
        print("saved %d kernels" % len(fp))
    elif "--compare" in args:
        old = json.load(open(args[args.index("--compare") + 1]))
        bad = 0
        for k, v in sorted(old.items()):
            if flt not in k:
                continue
            nk = k
            for a, b in maps:
                nk = nk.replace(a, b)
            if nk not in fp:
                print("MISSING  ", nk); bad += 1
            elif fp[nk] != v:
                print("DIFFERENT", nk, v, fp[nk]); bad += 1
        print("%d kernels compared, %d differ" % (len([k for k in old if flt in k]), bad))
        sys.exit(1 if bad else 0)
    else:
        for k, v in sorted(fp.items()):
            print(v[1], v[0], k)


if __name__ == "__main__":
    main()
