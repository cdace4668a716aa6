"""YAML-1.2 (core schema) loader on top of PyYAML.

Cantera mechanism files are YAML 1.2: ``NO`` is a species name (not a boolean) and ``1e13`` is a
float.  PyYAML implements YAML 1.1, so the implicit resolvers are replaced here.  The reference
front-end gets the same behaviour from ruamel.yaml (reference kinetix/core/mechanism.py:9,20).
"""
import re

import yaml

_BOOL = re.compile(r'^(?:true|True|TRUE|false|False|FALSE)$')
_NULL = re.compile(r'^(?:~|null|Null|NULL|)$')
_INT = re.compile(r'^(?:[-+]?[0-9]+|0o[0-7]+|0x[0-9a-fA-F]+)$')
_FLOAT = re.compile(r'^(?:[-+]?(?:\.[0-9]+|[0-9]+(?:\.[0-9]*)?)(?:[eE][-+]?[0-9]+)?'
                    r'|[-+]?\.(?:inf|Inf|INF)|\.(?:nan|NaN|NAN))$')


class Core12Loader(yaml.SafeLoader):
    """SafeLoader restricted to the YAML 1.2 core-schema scalar forms."""


Core12Loader.yaml_implicit_resolvers = {}
Core12Loader.add_implicit_resolver('tag:yaml.org,2002:bool', _BOOL, list('tTfF'))
Core12Loader.add_implicit_resolver('tag:yaml.org,2002:null', _NULL, ['~', 'n', 'N', ''])
Core12Loader.add_implicit_resolver('tag:yaml.org,2002:int', _INT, list('-+0123456789'))
Core12Loader.add_implicit_resolver('tag:yaml.org,2002:float', _FLOAT, list('-+0123456789.'))


def _construct_float(loader, node):
    text = loader.construct_scalar(node).lower()
    if text.endswith('.inf'):
        return float('-inf') if text.startswith('-') else float('inf')
    if text.endswith('.nan'):
        return float('nan')
    return float(text)


def _construct_int(loader, node):
    text = loader.construct_scalar(node)
    if text.startswith('0o'):
        return int(text[2:], 8)
    if text.startswith('0x'):
        return int(text[2:], 16)
    return int(text)


Core12Loader.add_constructor('tag:yaml.org,2002:float', _construct_float)
Core12Loader.add_constructor('tag:yaml.org,2002:int', _construct_int)


def load_file(path):
    with open(path) as fh:
        return yaml.load(fh, Loader=Core12Loader)
