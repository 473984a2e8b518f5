#!/bin/sh
# Pins the oracle to the real reference where a JRE exists: see tools/pin_reference.py.
exec python "$(dirname "$0")/pin_reference.py" "$@"
