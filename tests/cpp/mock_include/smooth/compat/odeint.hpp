#pragma once
// TEST INFRASTRUCTURE: smooth/compat/odeint.hpp makes Lie groups usable as odeint states; R^n vectors need nothing.
