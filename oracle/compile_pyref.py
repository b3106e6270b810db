#!/usr/bin/env python
"""Compile the reference's Python modules to sourceless bytecode under oracle/_ref/pyref/
(TEST INFRASTRUCTURE; runs only where /root/reference exists, the outputs travel to the GPU box
like the compiled reference C).  No reference source is copied: the .pyc files are build
products of the files where they lie (written as `name.bc`: snapshot tools tend to drop `*.pyc`),
imported through the small finder in oracle/pyref_loader.py.  tests/test_reference_stack_gpu.py
imports the reference's own sampler stack from there and runs it on the product path.

    python oracle/compile_pyref.py [/root/reference]
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
OUT = os.path.join(HERE, '_ref', 'pyref')


def wanted():
    """every module of the reference's root and of its clustering package"""
    out = []
    for sub in ('', 'clustering'):
        d = os.path.join(REF, sub)
        for name in sorted(os.listdir(d)):
            if name.endswith('.py'):
                out.append(os.path.join(sub, name) if sub else name)
    return out


def main():
    if not os.path.isdir(REF):
        print('reference tree %s not present: keeping prebuilt %s (if any)' % (REF, OUT))
        return
    n = 0
    for rel in wanted():
        src = os.path.join(REF, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(OUT, rel[:-3] + '.bc')
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        try:
            py_compile.compile(src, cfile=dst, dfile=rel, doraise=True)
            n += 1
        except py_compile.PyCompileError as e:      # a module that is not Python 3: not on the path we run
            print('skipped %s: %s' % (rel, str(e).splitlines()[0]))
    print('compiled %d reference modules into %s' % (n, OUT))


if __name__ == '__main__':
    main()
