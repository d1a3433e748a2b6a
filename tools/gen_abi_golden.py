"""Compile a tiny C program against the REFERENCE header and record every enum value / struct size that is
part of the public ABI -> tests/golden/abi_golden.json.  tests/test_abi.py compiles the same program against
include/pixelforge.h and compares."""
import json, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def enum_names(header_text):
    names = []
    for body in re.findall(r"typedef\s+enum\s*\{(.*?)\}\s*\w+\s*;", header_text, flags=re.S):
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S); body = re.sub(r"//[^\n]*", "", body)
        body = re.sub(r"#\s*(ifndef|ifdef|if|endif|else)[^\n]*", "", body)
        for item in body.split(","):
            m = re.match(r"\s*([A-Za-z_][A-Za-z0-9_]*)", item)
            if m: names.append(m.group(1))
    return [n for n in names if n.startswith("PF_")]

def probe(include_dir, names, extra_defs=""):
    src = ['#include "pixelforge.h"', "#include <stdio.h>", "#include <stddef.h>", "int main(void){"]
    for n in names: src.append(f'printf("{n} %lld\\n", (long long){n});')
    for t in ("PFcolor", "PFframebuffer", "PFboolean", "PFsizei", "PFenum", "PFcontext", "PFtexture", "PFrenderlist"):
        src.append(f'printf("sizeof_{t} %zu\\n", sizeof({t}));')
    src.append('printf("offsetof_PFframebuffer_zbuffer %zu\\n", offsetof(PFframebuffer, zbuffer));')
    src.append("return 0;}")
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c"); open(c, "w").write("\n".join(src))
        subprocess.run(["gcc", "-std=gnu99", f"-I{include_dir}", c, "-o", os.path.join(d, "p")], check=True)
        out = subprocess.run([os.path.join(d, "p")], capture_output=True, text=True, check=True).stdout
    return {l.split()[0]: int(l.split()[1]) for l in out.splitlines()}

def prototypes(header_text):
    text = re.sub(r"/\*.*?\*/", "", header_text, flags=re.S); text = re.sub(r"//[^\n]*", "", text)
    protos = {}
    for m in re.finditer(r"PF_API\s+([^;{]*?)\b(pf[A-Z]\w*)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.groups()
        norm = lambda s: re.sub(r"\s+", " ", s).strip()
        types = [re.sub(r"\b[a-zA-Z_]\w*$", "", norm(a)).strip() if norm(a) != "void" else "void" for a in args.split(",")]
        protos[name] = [norm(ret)] + [re.sub(r"\s*\*\s*", "*", t) for t in types]
    return protos

if __name__ == "__main__":
    ref = "/root/reference/src"
    text = open(os.path.join(ref, "pixelforge.h")).read()
    names = enum_names(text)
    out = {"values": probe(ref, names), "prototypes": prototypes(text)}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "abi_golden.json"), "w"), indent=1, sort_keys=True)
    print(len(out["values"]), "values,", len(out["prototypes"]), "prototypes")
