"""TEST INFRASTRUCTURE ONLY -- recipe that compiles the one compiled piece of the reference on a path this repo covers: the
PCL loss CPU op (projects/WSL/wsl/layers/csrc/pcl_loss/pcl_loss_cpu.cpp), from the reference's sources where they lie, into
oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).  Nothing is copied from /root/reference.

    python oracle/build_ref.py          # needs /root/reference; a no-op returning None when it is absent
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CSRC = "/root/reference/projects/WSL/wsl/layers/csrc"
OUT = os.path.join(HERE, "_ref")
NAME = "wsl_pcl_ref"


def _find_built():
    if not os.path.isdir(OUT):
        return None
    for f in os.listdir(OUT):
        if f.startswith(NAME) and f.endswith(".so"):
            return os.path.join(OUT, f)
    return None


def load_pcl_ref(build=True):
    """The compiled reference op as a Python module with pcl_loss_forward / pcl_loss_backward, or None."""
    so = _find_built()
    if so is None and build and os.path.isdir(REF_CSRC):
        from torch.utils.cpp_extension import load

        os.makedirs(OUT, exist_ok=True)
        return load(name=NAME, sources=[os.path.join(HERE, "pcl_ref_binding.cpp"), os.path.join(REF_CSRC, "pcl_loss", "pcl_loss_cpu.cpp")],
                    extra_include_paths=[REF_CSRC], extra_cflags=["-O2", "-w"], build_directory=OUT, verbose=False)
    if so is None:
        return None
    import torch  # noqa: F401  (the extension links against libtorch)

    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    m = load_pcl_ref()
    print("oracle/_ref:", "built " + str(_find_built()) if m is not None else "reference sources absent, nothing built")
    sys.exit(0)
