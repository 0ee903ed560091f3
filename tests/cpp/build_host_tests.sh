#!/usr/bin/env bash
# Builds tests/cpp/host_api_test against the in-tree engine library.  g++ only (the host mirror is header-only C++11/14).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"; ROOT="$(cd "$HERE/../.." && pwd)"
g++ -O2 -std=c++14 -Wall -I"$ROOT/include" -o "$HERE/host_api_test" "$HERE/host_api_test.cpp" \
    -L"$ROOT/openkite_b200" -lkite_b200 -Wl,-rpath,"$ROOT/openkite_b200"
