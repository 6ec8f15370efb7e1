#pragma once
#include <nanobind/nanobind.h>
