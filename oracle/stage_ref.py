"""ORACLE tooling (test infrastructure): stage the UNMODIFIED reference sources the denoiser hot path
and its callers need into oracle/_ref/reference, so that the real reference travels to the GPU box
(/root/reference does not exist there) as the checker and as the CPU arm of bench.py.

    python -m oracle.stage_ref            (also run by __graft_entry__.build() when /root/reference exists)

oracle/_ref/ is git-ignored (no reference source ever enters the history) but NOT gpurun-ignored.
The files are byte-for-byte copies; `MANIFEST.json` next to them records sha256 per file and the
reference commit, and `verify()` re-checks the staged copies against it.  Only what imports on the
sampling path is staged: models/, diffusion/, cond_gen/, sampling.py, mix_dpm_solver.py, utils.py,
configs/, datasets/datasets_config.py (pure-Python histogram tables).  Nothing under oracle/_ref is
imported by the product package (jodo_b200/): only tests/, smoke() and bench.py's reference arm do.
"""
import glob
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get('JODO_REFERENCE_ROOT', '/root/reference')
DST = os.path.join(HERE, '_ref', 'reference')
PATTERNS = ['models/*.py', 'diffusion/*.py', 'cond_gen/*.py', 'sampling.py', 'mix_dpm_solver.py', 'utils.py',
            'configs/*.py', 'datasets/datasets_config.py']


def _sha(path):
    with open(path, 'rb') as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(src=SRC, dst=DST):
    """Copy the listed files (unmodified) and write the manifest.  Returns the number of files."""
    if not os.path.isdir(os.path.join(src, 'models')):
        raise RuntimeError(f'reference tree not found at {src}')
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    files = {}
    for pat in PATTERNS:
        for p in sorted(glob.glob(os.path.join(src, pat))):
            rel = os.path.relpath(p, src)
            out = os.path.join(dst, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(p, out)
            files[rel] = _sha(out)
    commit = None
    sub = os.path.join(src, '.SUBMODULES.json')
    if os.path.exists(sub):
        try:
            with open(sub) as f:
                commit = json.load(f).get('commit')
        except (ValueError, OSError):
            commit = None
    with open(os.path.join(dst, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': src, 'commit': commit, 'files': files}, f, indent=1, sort_keys=True)
    return len(files)


def verify(dst=DST):
    """True when every staged file still matches the manifest (i.e. is the unmodified reference file)."""
    man = os.path.join(dst, 'MANIFEST.json')
    if not os.path.exists(man):
        return False
    with open(man) as f:
        files = json.load(f)['files']
    return all(os.path.exists(os.path.join(dst, rel)) and _sha(os.path.join(dst, rel)) == h for rel, h in files.items())


if __name__ == '__main__':
    n = stage()
    print(f'staged {n} reference files into {DST}; verify = {verify()}')
