// K5 — dense weighted least-squares voting of CoordLSVotingWeighted.calc (implemented below).
#pragma once
#include "common.cuh"
namespace casa {}
