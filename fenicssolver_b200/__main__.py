import sys

from .main import main

if len(sys.argv) < 2:
    print("usage: python -m fenicssolver_b200 case.json")
    sys.exit(2)
main(sys.argv[1])
