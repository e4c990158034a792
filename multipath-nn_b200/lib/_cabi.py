"""ctypes binding of libmpnn_sm100.so.  Prototypes are parsed from
include/mpnn.h so the Python side cannot drift from the C ABI.  There is no
fallback: if the library is missing, importing the engine fails loudly.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libmpnn_sm100.so')
HEADER = os.path.normpath(os.path.join(HERE, '..', '..', 'include', 'mpnn.h'))

_SCALARS = {'int': ctypes.c_int, 'float': ctypes.c_float, 'double': ctypes.c_double, 'long': ctypes.c_longlong,
            'unsigned': ctypes.c_uint}


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes], [argnames])} for every prototype."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
    src = re.sub(r'//[^\n]*', ' ', src)
    protos = {}
    for m in re.finditer(r'\b(int|const char\s*\*)\s+(mpnn_\w+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = ctypes.c_char_p if 'char' in ret else ctypes.c_int
        argtypes, argnames = [], []
        args = ' '.join(args.split())
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                argnames.append(re.findall(r'\w+', a)[-1])
                if '*' in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    argtypes.append(_SCALARS[re.findall(r'\w+', a)[-2]])
        protos[name] = (restype, argtypes, argnames)
    return protos


class MpnnError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                'libmpnn_sm100.so is not built (%s). Run `python __graft_entry__.py build` '
                '(nvcc, sm_100a). The product has no CPU fallback.' % LIB_PATH)
        self.dll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, argtypes, _) in self.protos.items():
            fn = getattr(self.dll, name)        # AttributeError if the symbol is missing
            fn.restype = restype
            fn.argtypes = argtypes
        self.launches = 0

    def __getattr__(self, name):
        if not name.startswith('mpnn_'):
            name = 'mpnn_' + name
        fn = getattr(self.dll, name)
        restype = self.protos[name][0]
        if restype is not ctypes.c_int or name in ('mpnn_version', 'mpnn_has_umma', 'mpnn_nccl_version'):
            return fn

        def call(*args):
            rc = fn(*args)
            if rc != 0:
                raise MpnnError('%s failed (%d): %s' % (name, rc, self.dll.mpnn_last_error().decode()))
            self.launches += 1
            return rc
        self.__dict__[name[5:]] = call
        self.__dict__[name] = call
        return call


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
