// all.h — every built-in single-source transition header.
#pragma once
#include "gol.h"
#include "hk.h"
#include "market.h"
#include "predator.h"
#include "sir.h"
#include "testkit.h"
