"""Option containers and temp-file helpers (host glue; mirrors buffalo/misc/_aux.py's public names).

``Option`` is a dict whose keys are also attributes (missing attribute -> None, nested dicts wrapped),
``InputOptions`` is the default/validation protocol of every ``*Option`` class: an option is valid when
every default key is present with the default's type (buffalo/misc/_aux.py:71-80).
"""
import abc
import atexit
import json
import os
import subprocess
import tempfile

_tracked_tmp = []


class Option(dict):
    """dict with attribute access; accepts dicts or JSON file paths (buffalo/misc/_aux.py:16-60)."""

    def __init__(self, *sources, **kwargs):
        merged = {}
        for src in sources:
            if not isinstance(src, dict):
                with open(src) as fin:
                    src = json.load(fin)
            merged.update(src)
        merged.update(kwargs)
        super().__init__()
        for k, v in merged.items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        return Option(v) if isinstance(v, dict) and not isinstance(v, Option) else v

    def __setitem__(self, key, value):
        value = Option._wrap(value)
        dict.__setitem__(self, key, value)
        self.__dict__[key] = value

    def __getattr__(self, name):  # only reached when the attribute is missing
        return self.get(name)

    def __setattr__(self, key, value):
        self[key] = value

    def __delitem__(self, key):
        dict.__delitem__(self, key)
        self.__dict__.pop(key, None)

    __delattr__ = __delitem__

    def update(self, *a, **kw):
        for k, v in dict(*a, **kw).items():
            self[k] = v

    def __getstate__(self):
        return dict(self)

    def __setstate__(self, state):
        for k, v in state.items():
            self[k] = v

    def __reduce__(self):
        return (Option, (dict(self),))


class InputOptions(abc.ABC):
    def __init__(self, *args, **kwargs):
        pass

    @abc.abstractmethod
    def get_default_option(self):
        raise NotImplementedError

    def is_valid_option(self, opt):
        defaults = self.get_default_option()
        for key, dflt in defaults.items():
            if key not in opt:
                raise RuntimeError("{} not exists on Option".format(key))
            if not isinstance(opt.get(key), type(dflt)):
                raise RuntimeError("Invalid type for {}, {} expected. ".format(key, type(dflt)))
        return True

    def create_temporary_option_from_dict(self, opt):
        """The native side takes the PATH of a JSON option file (buffalo/algo/base.py:18-24)."""
        fd, path = tempfile.mkstemp(dir=opt.get("tmp_dir", "/tmp/"), text=True)
        with os.fdopen(fd, "w") as fout:
            fout.write(json.dumps(opt))
        _tracked_tmp.append(path)
        return path


def get_temporary_file(root="/tmp/", write_mode="w"):
    fd, path = tempfile.mkstemp(dir=root)
    os.close(fd)
    _tracked_tmp.append(path)
    return path


def register_cleanup_file(path):
    _tracked_tmp.append(path)


def copy_to_temporary_file(source_path, ignore_lines=0, chunk_size=8192, binary=False):
    path = get_temporary_file()
    with open(source_path, "rb" if binary else "r") as fin, open(path, "wb" if binary else "w") as fout:
        for _ in range(ignore_lines):
            fin.readline()
        while True:
            chunk = fin.read(chunk_size)
            if not chunk:
                break
            fout.write(chunk)
    return path


def psort(path, parallel=-1, field_seperator=" ", key=1, tmp_dir="/tmp/", buffer_mb=1024, output=None):
    """GNU sort wrapper kept for API compatibility (buffalo/misc/_aux.py:115-137)."""
    cmd = ["sort", "-n", "-s", "-t", field_seperator, "-k", str(key), "-T", tmp_dir, "-S", "%sM" % buffer_mb,
           "-o", output or path, path]
    if parallel != 0:
        cmd[3:3] = ["--parallel", str(os.cpu_count() if parallel < 0 else parallel)]
    subprocess.check_output(cmd, stderr=subprocess.STDOUT, env={"LC_ALL": "C"})


@atexit.register
def _remove_tracked_tmp():
    for path in _tracked_tmp:
        try:
            if os.path.isfile(path):
                os.remove(path)
        except OSError:
            pass
