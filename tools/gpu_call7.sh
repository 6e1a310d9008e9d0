#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 --tb=short 2>&1 | grep -v Warning | tail -25
