"""YAML -> attribute-dict config with the semantics of utils/detzero_utils/config_utils.py:24-94
(``_BASE_CONFIG_`` include, ``--set KEY VAL`` typed overrides).  ``easydict`` is not in the image, so a small
attribute dict stands in for EasyDict."""
import os
from ast import literal_eval

import yaml


class AttrDict(dict):
    """dict with attribute access, recursively (EasyDict stand-in)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            return AttrDict(v)
        if isinstance(v, (list, tuple)):
            return type(v)(AttrDict._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, AttrDict._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def update(self, other=None, **kw):
        for k, v in dict(other or {}, **kw).items():
            self[k] = v


def merge_new_config(config, new_config, base_dir=None):
    """config_utils.py:59-76.  The reference resolves ``_BASE_CONFIG_`` relative to the CWD (scripts run from
    tools/); here the including file's directory tree is searched too, so configs work from anywhere."""
    if '_BASE_CONFIG_' in new_config:
        path = new_config['_BASE_CONFIG_']
        cands = [path]
        d = base_dir
        while d and d != os.path.dirname(d):
            cands.append(os.path.join(d, path))
            d = os.path.dirname(d)
        for c in cands:
            if os.path.exists(c):
                with open(c, 'r') as f:
                    config.update(AttrDict(yaml.safe_load(f)))
                break
        else:
            raise FileNotFoundError(path)
    for key, val in new_config.items():
        if not isinstance(val, dict):
            config[key] = val
            continue
        if key not in config:
            config[key] = AttrDict()
        merge_new_config(config[key], val, base_dir)
    return config


def cfg_from_yaml_file(cfg_file, config):
    with open(cfg_file, 'r') as f:
        new_config = yaml.safe_load(f)
    merge_new_config(config=config, new_config=new_config, base_dir=os.path.dirname(os.path.abspath(cfg_file)))
    return config


def cfg_from_list(cfg_list, config):
    """``--set KEY VAL ...`` overrides (config_utils.py:24-56)."""
    assert len(cfg_list) % 2 == 0
    for k, v in zip(cfg_list[0::2], cfg_list[1::2]):
        keys = k.split('.')
        d = config
        for sub in keys[:-1]:
            assert sub in d, 'NotFoundKey: %s' % sub
            d = d[sub]
        sub = keys[-1]
        assert sub in d, 'NotFoundKey: %s' % sub
        try:
            value = literal_eval(v)
        except Exception:
            value = v
        if type(value) != type(d[sub]) and isinstance(d[sub], AttrDict):
            for src in value.split(','):
                ck, cv = src.split(':')
                d[sub][ck] = type(d[sub][ck])(cv)
        elif type(value) != type(d[sub]) and isinstance(d[sub], list):
            items = list(value) if isinstance(value, (tuple, list)) else value.split(',')
            d[sub] = [type(d[sub][0])(x) for x in items]
        else:
            assert type(value) == type(d[sub]), 'type %s does not match %s' % (type(value), type(d[sub]))
            d[sub] = value


cfg = AttrDict()
cfg.LOCAL_RANK = 0
