"""Compile the oracle's C restatement (gcc).  `oracle/_ref` does not exist for this reference: its
native code is Objective-C++ against Apple CoreML (coreml/coreml.mm) and cannot be built on Linux;
the reference's runnable path is Python and is exercised by tests/golden/make_golden.py instead."""
from oracle import timing

if __name__ == "__main__":
    print(timing.build(force=True))
