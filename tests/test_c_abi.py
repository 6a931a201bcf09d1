"""The drop-in boundary: libsiesta_gpu.so loads without a GPU and exports exactly the functions include/siesta_gpu.h
declares (no compute calls here); calls that need a device fail loudly instead of falling back to anything."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "siesta_gpu.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)   # comments
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)      # preprocessor lines
    return sorted(set(re.findall(r"\b(siesta_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_function():
    from sequencedetectionqueryexecutor_b200 import _lib
    names = declared_functions()
    assert len(names) >= 30
    L = C.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in siesta_gpu.h but not exported: {missing}"
    # and the binding knows every one of them
    unbound = [n for n in names if n not in _lib.EXPORTS]
    assert not unbound, f"declared but absent from the Python binding's EXPORTS: {unbound}"


def test_exported_siesta_symbols_are_all_declared():
    from sequencedetectionqueryexecutor_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted({ln.split()[-1] for ln in out.splitlines() if re.search(r"\sT\ssiesta_[a-z0-9_]+$", ln)})
    undeclared = [n for n in exported if n not in declared_functions()]
    assert not undeclared, f"exported extern \"C\" functions missing from siesta_gpu.h: {undeclared}"


def test_no_cpu_fallback_without_a_device():
    """Without a usable sm_100 device siesta_init must fail (SIESTA_E_CUDA) - the product has no CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the failure path of siesta_init cannot be observed")
    from sequencedetectionqueryexecutor_b200 import _abi, _lib
    h = C.c_void_p()
    rc = _lib.lib().siesta_init(0, C.byref(h))
    assert rc == _abi.E_CUDA and not h.value
    assert b"" != _lib.lib().siesta_last_error()


def test_jni_shim_implements_every_native_method_of_the_java_class():
    """jni/siesta_gpu_jni.c (compiled against jni/stub/jni.h: no JDK in this image) defines one
    Java_..._GpuNative_<name> per `static native` method of jni/java/.../GpuNative.java, and links against the library."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "jni"), "-s"])
    java = open(os.path.join(ROOT, "jni", "java", "com", "datalab", "siesta", "queryprocessor", "SaseConnection", "GpuNative.java")).read()
    natives = sorted(set(re.findall(r"static native [\w\[\]]+ (\w+)\(", java)))
    assert len(natives) >= 10
    so = os.path.join(ROOT, "jni", "_build", "libsiesta_gpu_jni.so")
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    prefix = "Java_com_datalab_siesta_queryprocessor_SaseConnection_GpuNative_"
    defined = sorted(ln.split()[-1][len(prefix):] for ln in out.splitlines() if prefix in ln)
    assert defined == natives
    needed = subprocess.run(["readelf", "-d", so], capture_output=True, text=True).stdout
    assert "libsiesta_gpu.so" in needed
