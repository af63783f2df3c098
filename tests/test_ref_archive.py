"""oracle/build_ref.py: the unmodified reference travels to the GPU box as ONE archive (a build artefact under the
git-ignored oracle/_ref/), is verified against its manifest when unpacked, and is what oracle/ref_import.py falls back
to when /root/reference is absent."""
import hashlib
import json
import os
import subprocess
import sys
import tarfile

import pytest

from oracle import build_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(build_ref.build() is None, reason="neither the reference sources nor a built archive on this machine")


def test_archive_members_are_byte_identical_to_the_manifest(tmp_path):
    files = json.load(open(build_ref.MANIFEST))["files"]
    assert set(files) == set(build_ref.FILES)
    with tarfile.open(build_ref.ARCHIVE, "r:gz") as tar:
        names = [m.name for m in tar.getmembers()]
        assert sorted(names) == sorted(files)
        for m in tar.getmembers():
            assert hashlib.sha256(tar.extractfile(m).read()).hexdigest() == files[m.name]
    if build_ref.source_available():          # build container: the manifest is the reference's own files
        for f, h in files.items():
            assert build_ref._sha(os.path.join(build_ref.SRC_ROOT, f)) == h
    # no reference source exists as a plain file of the repository
    assert sorted(os.listdir(build_ref.DST_ROOT)) == ["MANIFEST.json", "reference_src.tar.gz"]
    d = build_ref.unpack(str(tmp_path / "ref"))
    assert os.path.isfile(os.path.join(d, "core", "raycasters.py")) and os.path.isfile(os.path.join(d, "configs", "surreal", "surreal.txt"))


def test_import_falls_back_to_the_archive_when_the_sources_are_absent():
    code = (
        "import os, sys; sys.path.insert(0, %r)\n"
        "real = os.path.isfile\n"
        "os.path.isfile = lambda p: False if str(p).startswith('/root/reference') else real(p)\n"
        "from oracle import ref_import as ri\n"
        "assert ri.reference_available() and not ri.REF_ROOT.startswith('/root/reference'), ri.REF_ROOT\n"
        "core = ri.import_reference()\n"
        "import core.raycasters as r, core.trainer as t, core.pose_opt as p\n"
        "assert r.__file__.startswith(ri.REF_ROOT) and not r.__file__.startswith(%r)\n"
        "print('ok')\n" % (ROOT, ROOT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env={k: v for k, v in os.environ.items() if k != "ANERF_REFERENCE_ROOT"})
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-1500:]
