"""``dict_from_file``: the hand-rolled ``key = value`` config reader Feature_Grid_Inference.py needs
(reference: visualization/pltUtils.py:24-63).  The plotting helpers of that file are out of scope."""
from __future__ import annotations


def _parse_value(text):
    for cast in (int, float):
        try:
            return cast(text)
        except ValueError:
            pass
    if ',' in text:
        items = text.replace('[', '').replace(']', '').split(',')
        try:
            return [int(x) for x in items]
        except ValueError:
            return [float(x) for x in items]
    if text in ('True', 'False'):
        return bool(text)  # sic: the reference maps both spellings to True
    return text


def dict_from_file(filename):
    out = {}
    with open(filename, 'r') as f:
        for line in f:
            parts = line.replace(' ', '').replace('\n', '').split('=')
            if len(parts) > 1:
                out[parts[0]] = _parse_value(parts[1])
    return out
