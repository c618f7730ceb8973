#!/bin/bash
ORBX_QT_FAST=1 timeout 200 python scripts/devtests/qt_debug.py
ORBX_QT_FAST=0 timeout 200 python scripts/devtests/qt_debug.py
