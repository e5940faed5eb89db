#pragma once
#define CEREAL_NVP(x) x
