"""Stand-in for ConfigArgParse 1.5.3 (absent from this image; the reference's CLI scripts import it,
Feature_Grid_Training.py:5, Feature_Grid_Inference.py:28).  TEST INFRASTRUCTURE: only what those two scripts use --
``ArgumentParser`` with ``add_argument(..., is_config_file=True)`` and ``key = value`` config files whose entries act
as defaults that the command line may override; ``key =`` and ``key = ''`` give the empty string, ``[a, b]`` a list."""
import argparse


class ArgumentParser(argparse.ArgumentParser):

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._config_dests = []

    def add_argument(self, *a, **kw):
        is_cfg = kw.pop('is_config_file', False)
        act = super().add_argument(*a, **kw)
        if is_cfg:
            self._config_dests.append(act)
        return act

    def parse_known_args(self, args=None, namespace=None):
        import sys
        argv = list(sys.argv[1:] if args is None else args)
        extra = []
        for act in self._config_dests:
            for flag in act.option_strings:
                for i, tok in enumerate(argv):
                    path = None
                    if tok == flag and i + 1 < len(argv):
                        path = argv[i + 1]
                    elif tok.startswith(flag + '='):
                        path = tok.split('=', 1)[1]
                    if path:
                        extra += self._read(path)
        # config entries first, so that explicit command-line flags win
        return super().parse_known_args(extra + argv, namespace)

    def _read(self, path):
        known = {a.dest: a for a in self._actions}
        out = []
        with open(path) as f:
            for line in f:
                line = line.strip()
                if not line or line[0] in '#;' or '=' not in line:
                    continue
                k, v = [s.strip() for s in line.split('=', 1)]
                if k not in known:
                    continue
                flag = known[k].option_strings[0]
                if len(v) >= 2 and v[0] == v[-1] and v[0] in '\'"':
                    v = v[1:-1]
                if v.startswith('[') and v.endswith(']'):
                    out += [flag] + [s.strip() for s in v[1:-1].split(',') if s.strip()]
                else:
                    out.append('%s=%s' % (flag, v))
        return out
