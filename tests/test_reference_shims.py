"""CPU checks of the stand-ins used to compile the reference's own sources against our headers (oracle/ref_shim/user):
a wrong EXPECT_* in the googletest stand-in would turn into false passes or failures of oracle/_ref/ref_gtests."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_gtest_standin_semantics(tmp_path):
    """TEST registration, EXPECT_EQ on atomics, EXPECT_FLOAT_EQ = 4 ULP (sign-aware), EXPECT_NEAR, EXPECT_STREQ,
    ASSERT_* leaving the test body: one test that must pass, one that must record exactly four failures."""
    exe = str(tmp_path / "gt_self")
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "oracle", "ref_shim", "user"), "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "gtest_shim_selftest.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True).stdout.split("\n")
    assert out[0] == "Shim.Passes failures 0" and out[1] == "Shim.Fails failures 4", out


def test_reference_source_builds_are_declared():
    """the recipes that compile reference sources live under oracle/ and write into oracle/_ref/ only"""
    mk = open(os.path.join(ROOT, "oracle", "Makefile")).read()
    for target in ("_ref/libshim_refsrc1.so", "_ref/libshim_refsrc2.so", "_ref/ref_gtests", "_ref/ref_gpu_queries", "_ref/libvn_ref.so"):
        assert target in mk, target
    assert "oracle/_ref/" in open(os.path.join(ROOT, ".gitignore")).read()
    ign = os.path.join(ROOT, ".gpurunignore")
    if os.path.exists(ign):  # the binaries must travel to the GPU box: only the object files may be excluded
        assert [l.strip() for l in open(ign) if l.strip() and not l.startswith("#")] in ([], ["oracle/_ref/obj/"])
