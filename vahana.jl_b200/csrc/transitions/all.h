// all.h — every built-in single-source transition header.
#pragma once
#include "hk.h"
#include "testkit.h"
