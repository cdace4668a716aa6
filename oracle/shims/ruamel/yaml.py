"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for ``ruamel.yaml`` (absent in this image).

The reference generator imports ``from ruamel.yaml import YAML`` (reference
kinetix/core/mechanism.py:9) and only ever calls ``YAML().load(open(path))``
(mechanism.py:20).  This shim provides exactly that on top of PyYAML, but with
YAML-1.2 *core schema* scalar resolution, which is what ruamel implements:
  * booleans are only true/false (so the species name ``NO`` stays a string),
  * ``1e13`` / ``2.0E+06`` without a dot are floats,
  * no sexagesimal ints, no ``yes/no/on/off``.
It lets ``oracle/build_ref.py`` run the reference generator UNMODIFIED from
/root/reference.  Nothing in the product package imports this file.
"""
import re
import yaml as _pyyaml


class _Core12Loader(_pyyaml.SafeLoader):
    pass


# drop PyYAML's YAML-1.1 implicit resolvers, install the 1.2 core-schema ones
_Core12Loader.yaml_implicit_resolvers = {}
_Core12Loader.add_implicit_resolver(
    'tag:yaml.org,2002:bool', re.compile(r'^(?:true|True|TRUE|false|False|FALSE)$'), list('tTfF'))
_Core12Loader.add_implicit_resolver(
    'tag:yaml.org,2002:null', re.compile(r'^(?:~|null|Null|NULL|)$'), ['~', 'n', 'N', ''])
_Core12Loader.add_implicit_resolver(
    'tag:yaml.org,2002:int', re.compile(r'^(?:[-+]?[0-9]+|0o[0-7]+|0x[0-9a-fA-F]+)$'), list('-+0123456789'))
_Core12Loader.add_implicit_resolver(
    'tag:yaml.org,2002:float',
    re.compile(r'^(?:[-+]?(?:\.[0-9]+|[0-9]+(?:\.[0-9]*)?)(?:[eE][-+]?[0-9]+)?'
               r'|[-+]?\.(?:inf|Inf|INF)|\.(?:nan|NaN|NAN))$'),
    list('-+0123456789.'))


def _float(loader, node):
    s = loader.construct_scalar(node).lower()
    if s.endswith('.inf'):
        return float('-inf') if s.startswith('-') else float('inf')
    if s.endswith('.nan'):
        return float('nan')
    return float(s)


def _int(loader, node):
    s = loader.construct_scalar(node)
    if s.startswith('0o'):
        return int(s[2:], 8)
    if s.startswith('0x'):
        return int(s[2:], 16)
    return int(s)


_Core12Loader.add_constructor('tag:yaml.org,2002:float', _float)
_Core12Loader.add_constructor('tag:yaml.org,2002:int', _int)


class YAML:
    def __init__(self, *args, **kwargs):
        pass

    def load(self, stream):
        return _pyyaml.load(stream, Loader=_Core12Loader)
