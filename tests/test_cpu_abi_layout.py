"""The ctypes mirrors in osrl_b200/_lib.py against the C header itself: a C program that includes
include/osrl_b200.h (plain C, gcc -- the header is the boundary a non-Python host would compile against) prints
sizeof and the offset of every field; the ctypes Structures must agree field by field."""
import ctypes as C
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAIRS = {"osrl_config": "Config", "osrl_param_desc": "ParamDesc", "osrl_dataset_view": "DatasetView", "osrl_batch": "Batch",
         "osrl_noise": "Noise", "osrl_seq_batch": "SeqBatch", "osrl_seq_dataset_view": "SeqDatasetView"}


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_ctypes_structs_match_the_c_header(tmp_path):
    from osrl_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "osrl_b200.h")).read()
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "osrl_b200.h"', "int main(void) {"]
    fields = {}
    for cname in PAIRS:
        body = re.search(r"typedef struct " + cname + r" \{(.*?)\} " + cname + ";", hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            # "const float* a, *b" / "int32_t a_hidden[OSRL_MAX_HIDDEN]" / "float reward_scale, cost_scale"
            first, *rest = decl.split(",")
            names.append(re.sub(r"\[.*", "", first.split()[-1].lstrip("*")))
            names += [re.sub(r"\[.*", "", r.strip().lstrip("*")) for r in rest]
        fields[cname] = names
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for n in names:
            lines.append(f'  printf("{cname} {n} %zu\\n", offsetof({cname}, {n}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    got = {}
    for ln in out.splitlines():
        cname, key, val = ln.split()
        got.setdefault(cname, {})[key] = int(val)
    for cname, pyname in PAIRS.items():
        st = getattr(_lib, pyname)
        assert C.sizeof(st) == got[cname]["size"], cname
        py_fields = [f[0] for f in st._fields_]
        assert py_fields == fields[cname], (cname, py_fields, fields[cname])
        for n in py_fields:
            assert getattr(st, n).offset == got[cname][n], (cname, n)
