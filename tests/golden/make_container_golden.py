"""Have the REFERENCE ITSELF (oracle/_ref: ZFile.cpp + the tools library, compiled by oracle/build_ref.sh)
write the golden files of the file-format tests.  Run in the authoring container:

    python tests/golden/make_container_golden.py

  zfile_ref.bin   z_open_file_write / z_write_image x 6 / z_close_file on tests/container_cases.golden_movie()
  attrs_ref.bin   512 payload bytes, then the trailer attrs_set_times / attrs_set_global_attributes /
                  attrs_set_frame_attributes / attrs_close append (tests/container_cases.golden_attrs())
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import container as oc  # noqa: E402
from tests import container_cases as cc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    z = os.path.join(HERE, "zfile_ref.bin")
    n = oc.ref_write_zfile(z, cc.golden_movie(), cc.golden_times(), rate=50, clevel=2)
    print("wrote", z, os.path.getsize(z), "bytes (image data", n, ")")
    a = os.path.join(HERE, "attrs_ref.bin")
    g, frames, payload = cc.golden_attrs()
    with open(a, "wb") as f:
        f.write(payload)
    oc.ref_write_attrs(a, g, frames, cc.golden_times())
    print("wrote", a, os.path.getsize(a), "bytes")


if __name__ == "__main__":
    main()
