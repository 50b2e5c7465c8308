#pragma once
#include "ref_vkl_stub.h"
