// oracle shim: intentionally empty (see ../core/core.hpp)
#pragma once
