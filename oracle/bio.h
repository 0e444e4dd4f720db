// ORACLE — TEST INFRASTRUCTURE ONLY.  (filled in below)
#pragma once
