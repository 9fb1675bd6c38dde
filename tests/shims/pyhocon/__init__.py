"""TEST SHIM — a small stand-in for `pyhocon` (not installed in this image; no network), enough for the reference's
confs/*.conf and the ConfigTree calls exp_runner.py / models/dataset.py make: ConfigFactory.parse_string, dotted
`conf["a.b"]`, get / get_int / get_float / get_bool / get_string / get_list with defaults, put, `in`, and `**conf["x"]`.
Grammar subset: `key = value`, `key { ... }`, quoted keys, `[a, b]` lists, `#` / `//` comments, optional trailing commas.
Test infrastructure only: nothing under fmov_pose_b200/ imports it."""
import re
from collections import OrderedDict

_UNSET = object()


class ConfigMissingException(KeyError):
    pass


class ConfigException(Exception):
    pass


class ConfigTree(OrderedDict):
    def _walk(self, key, create=False):
        parts = key.split(".") if isinstance(key, str) else [key]
        node = self
        for p in parts[:-1]:
            if not (OrderedDict.__contains__(node, p) and isinstance(OrderedDict.__getitem__(node, p), ConfigTree)):
                if not create:
                    raise ConfigMissingException(f"No configuration setting found for key {key}")
                OrderedDict.__setitem__(node, p, ConfigTree())
            node = OrderedDict.__getitem__(node, p)
        return node, parts[-1]

    def __getitem__(self, key):
        node, last = self._walk(key)
        if not OrderedDict.__contains__(node, last):
            raise ConfigMissingException(f"No configuration setting found for key {key}")
        return OrderedDict.__getitem__(node, last)

    def __contains__(self, key):
        try:
            self[key]
            return True
        except ConfigMissingException:
            return False

    def put(self, key, value, append=False):
        node, last = self._walk(key, create=True)
        OrderedDict.__setitem__(node, last, value)

    def get(self, key, default=_UNSET):
        try:
            return self[key]
        except ConfigMissingException:
            if default is _UNSET:
                raise
            return default

    def _typed(self, key, default, conv):
        v = self.get(key, default)
        if v is default and default is not _UNSET:
            return v
        return conv(v)

    def get_int(self, key, default=_UNSET):
        return self._typed(key, default, lambda v: int(v))

    def get_float(self, key, default=_UNSET):
        return self._typed(key, default, lambda v: float(v))

    def get_string(self, key, default=_UNSET):
        return self._typed(key, default, lambda v: str(v))

    def get_list(self, key, default=_UNSET):
        return self._typed(key, default, lambda v: list(v))

    def get_config(self, key, default=_UNSET):
        return self.get(key, default)

    @staticmethod
    def _bool(v):
        if isinstance(v, str):
            if v.lower() in ("true", "yes", "on"):
                return True
            if v.lower() in ("false", "no", "off"):
                return False
            raise ConfigException(f"not a boolean: {v!r}")
        return bool(v)

    def get_bool(self, key, default=_UNSET):
        return self._typed(key, default, self._bool)


_TOKEN = re.compile(r'\s*(?:(#|//)[^\n]*|("(?:[^"\\]|\\.)*")|([{}\[\],=:\n])|([^\s{}\[\],=:#"]+))', re.S)


def _tokens(text):
    pos, out = 0, []
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            if text[pos:].strip() == "":
                break
            raise ConfigException(f"cannot parse near {text[pos:pos + 30]!r}")
        pos = m.end()
        if m.group(1):
            continue
        if m.group(2) is not None:
            out.append(("str", m.group(2)[1:-1]))
        elif m.group(3) is not None:
            out.append(("sym", m.group(3)))
        else:
            out.append(("word", m.group(4)))
    return out


def _scalar(kind, word):
    if kind == "str":
        return word
    lw = word.lower()
    if lw in ("true", "false"):
        return lw == "true"
    if lw == "null":
        return None
    try:
        return int(word)
    except ValueError:
        pass
    try:
        return float(word)
    except ValueError:
        return word


def _skip_nl(toks, i):
    while i < len(toks) and toks[i] == ("sym", "\n"):
        i += 1
    return i


def _parse_value(toks, i):
    kind, w = toks[i]
    if (kind, w) == ("sym", "{"):
        return _parse_object(toks, i + 1, closing=True)
    if (kind, w) == ("sym", "["):
        items, i = [], i + 1
        while True:
            i = _skip_nl(toks, i)
            if toks[i] == ("sym", "]"):
                return items, i + 1
            if toks[i] == ("sym", ","):
                i += 1
                continue
            v, i = _parse_value(toks, i)
            items.append(v)
    # unquoted scalars may consist of several adjacent words on one line (e.g. paths are single words here)
    return _scalar(kind, w), i + 1


def _parse_object(toks, i, closing):
    tree = ConfigTree()
    while True:
        i = _skip_nl(toks, i)
        if i >= len(toks):
            if closing:
                raise ConfigException("missing '}'")
            return tree, i
        if toks[i] == ("sym", "}"):
            if not closing:
                raise ConfigException("unexpected '}'")
            return tree, i + 1
        if toks[i] == ("sym", ","):
            i += 1
            continue
        kind, key = toks[i]
        if kind == "sym":
            raise ConfigException(f"unexpected {key!r}")
        i += 1
        if toks[i] == ("sym", "{"):
            val, i = _parse_object(toks, i + 1, closing=True)
            if key in tree and isinstance(tree[key], ConfigTree):          # HOCON merges repeated objects
                tree[key].update(val)
                continue
        elif toks[i][0] == "sym" and toks[i][1] in "=:":
            val, i = _parse_value(toks, _skip_nl(toks, i + 1))
        else:
            raise ConfigException(f"expected '=' or '{{' after key {key!r}")
        tree.put(key, val)


class ConfigFactory:
    @staticmethod
    def parse_string(text, **_):
        tree, _i = _parse_object(_tokens(text), 0, closing=False)
        return tree

    @staticmethod
    def parse_file(path, **_):
        with open(path) as fh:
            return ConfigFactory.parse_string(fh.read())
