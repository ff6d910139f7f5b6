"""Pair / image list parsing -- same file formats and key construction as imcui/hloc/utils/parsers.py:10-59."""
from collections import defaultdict
from pathlib import Path


def parse_image_list(path):
    """one image name per line (first token); blank lines and '#' comments skipped (parsers.py:10-31, without intrinsics)"""
    names = []
    for line in Path(path).read_text().split("\n"):
        line = line.strip()
        if line and not line.startswith("#"):
            names.append(line.split()[0])
    assert len(names) > 0, path
    return names


def parse_image_lists(paths):
    """`paths` may contain a glob in its last component (parsers.py:34-41)"""
    paths = Path(paths)
    files = sorted(paths.parent.glob(paths.name))
    assert len(files) > 0, paths
    return [n for f in files for n in parse_image_list(f)]


def parse_retrieval(path):
    """'query reference' per line -> {query: [references...]} in file order (parsers.py:44-52)"""
    out = defaultdict(list)
    for line in Path(path).read_text().rstrip("\n").split("\n"):
        if line:
            q, r = line.split()
            out[q].append(r)
    return dict(out)


def names_to_pair(name0, name1, separator="/"):
    """group key of a pair: '/' inside image names becomes '-' (parsers.py:54-55)"""
    return separator.join((name0.replace("/", "-"), name1.replace("/", "-")))


def names_to_pair_old(name0, name1):
    return names_to_pair(name0, name1, separator="_")
