// mock: see tests/native/dealii_mock/dealii_mock.h
#include "../../dealii_mock.h"
